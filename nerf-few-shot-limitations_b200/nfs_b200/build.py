"""In-tree build of the sm_100a C-ABI library (include/nfs_b200.h).

`python -m nfs_b200.build` (or `__graft_entry__.build()`) runs nvcc on every
csrc/*.cu and links one shared object next to this file.  nvcc cross-compiles
without a GPU; the resulting .so travels with the repo snapshot to the B200 box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(HERE)
CSRC = os.path.join(PKG_ROOT, "csrc")
LIB_PATH = os.path.join(HERE, "libnfs_b200.so")



def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(PKG_ROOT), "include", "nfs_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


VARIANTS = {"exactdiv": ["-DNFS_K1_EXACT_DIV"], "fastdiv": ["-DNFS_K1_FAST_DIV"], "devtools": ["-DNFS_DEVTOOLS"],
            "noloadhint": ["-DNFS_NO_LOAD_HINT"], "nostorehint": ["-DNFS_NO_STORE_HINT"]}      # developer A/B builds: libnfs_b200_<variant>.so (NFS_B200_LIB selects it)


def build(force=False, verbose=False, variant=None):
    if variant is not None:
        return _build(True, verbose, os.path.join(HERE, "libnfs_b200_%s.so" % variant), "build_" + variant, VARIANTS[variant])
    if not force and not needs_build():
        return LIB_PATH
    return _build(force, verbose, LIB_PATH, "build", [])


def _build(force, verbose, lib_path, objdir_name, extra):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libnfs_b200.so")
    objs = []
    objdir = os.path.join(PKG_ROOT, objdir_name)
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        newest = max([os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)] + [os.path.getmtime(
            os.path.join(os.path.dirname(PKG_ROOT), "include", "nfs_b200.h"))])
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > newest:
            continue
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-Xcompiler", "-fPIC", "-c", src, "-o", obj] + list(extra)
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        if os.environ.get("NFS_DEVTOOLS", "0") not in ("", "0"):
            cmd.insert(1, "-DNFS_DEVTOOLS")          # bisection switches + timeline tracer in the fused MLP kernel
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out))
        if verbose and out:
            print(out)
    tmp = lib_path + ".tmp"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", tmp] + objs + ["-lcuda"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    os.replace(tmp, lib_path)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv,
                variant=sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None))
