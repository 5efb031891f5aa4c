"""ctypes binding of the C ABI declared in include/nfs_b200.h.

The shared object is built in-tree by nfs_b200/build.py.  There is no CPU path and
no fallback: if the library is missing, or an entry point returns non-zero, a
RuntimeError naming the kernel is raised (SURVEY.md section 8b "Error convention").
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NFS_B200_LIB") or os.path.join(_HERE, "libnfs_b200.so")     # NFS_B200_LIB: developer A/B builds
HEADER_PATH = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "nfs_b200.h")

_p = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int32
_f32 = ctypes.c_float

# name -> (restype, argtypes); must stay in sync with include/nfs_b200.h
# (tests/test_abi.py parses the header and checks both directions).
SIGNATURES = {
    "nfs_abi_version": (ctypes.c_int, []),
    "nfs_last_error_string": (ctypes.c_char_p, []),
    "nfs_launch_count": (ctypes.c_uint64, []),
    "nfs_set_debug_flags": (None, [_i32]),
    "nfs_set_debug_trace": (None, [_p]),
    "nfs_composite_fwd": (ctypes.c_int, [_p, _p, _p, _p, _p, _f32, _i64, _i32, _i32, _i32, _p, _p, _p, _p]),
    "nfs_composite_bwd": (ctypes.c_int, [_p, _p, _p, _p, _p, _f32, _p, _p, _p, _i64, _i32, _i32, _i32, _p, _p, _p]),
    "nfs_composite_bwd_dy": (ctypes.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _p, _i64, _p]),
    "nfs_composite_loss_fwd": (ctypes.c_int, [_p, _p, _p, _p, _p, _f32, _p, _p, _f32, _f32, _i64, _i32, _i32, _i32,
                                              _p, _p, _p, _p, _p, _p, _p]),
    "nfs_rays_generate": (ctypes.c_int, [_i32, _i32, _f32, _p, _i32, _p, _i64, _p, _p, _p, _p, _p]),
    "nfs_posenc_fwd": (ctypes.c_int, [_p, _p, _i64, _i32, _i32, _i32, _p, _p]),
    "nfs_sample_stratified": (ctypes.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i32, _p, _p, _p]),
    "nfs_project_gather": (ctypes.c_int, [_p, _p, _f32, _i32, _i32, _p, _i32, _i32, _i32, _i64, _p, _p, _p, _p, _p]),
    "nfs_linear_bf16": (ctypes.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _p, _i64, _p, _p]),
    "nfs_wgrad_bf16": (ctypes.c_int, [_p, _i64, _p, _i64, _i64, _i32, _i32, _i32, _i32, _p, _i64, _i64, _p, _i32, _p]),
    "nfs_wgrad_multi_bf16": (ctypes.c_int, [_p, _i32, _p]),
    "nfs_mlp_chain": (ctypes.c_int, [_p, _i64, _i32, _p, _p, _p, _p, _p, _i32, _p, _p, _i64, _p, _p, _p, _i64, _p, _i32, _p]),
    "nfs_mlp_backward_fused": (ctypes.c_int, [_p, _i64, _i32, _p, _p, _p, _p, _p, _i32, _p, _i64, _p, _p, _i64, _p, _i32, _p, _p,
                                              _i32, _p]),
    "nfs_pack_stack": (ctypes.c_int, [_p, _i32, _i32, _p, _p, _p, _p]),
    "nfs_pack_table": (ctypes.c_int, [_p, _i32, _i32, _p]),
    "nfs_scatter_add_table": (ctypes.c_int, [_p, _i32, _i64, _p]),
    "nfs_g3_operand": (ctypes.c_int, [_p, _p, ctypes.c_float, _i32, _i32, _p, _i32, _i32, _i32, _p, _i32, _i32, _i64, _i32,
                                      _i64, _p, _p]),
    "nfs_gate_scale_bf16": (ctypes.c_int, [_p, _i64, _p, _i64, _i32, _i32, _p, _i64, _p]),
    "nfs_gate_bwd_operand": (ctypes.c_int, [_p, _i64, _p, _p, _i64, _i64, _i32, _i32, _i32, _p, _p]),
    "nfs_bias_terms_bf16": (ctypes.c_int, [_p, _i32, _p, _p]),
    "nfs_mlp_chain_points": (ctypes.c_int, [_p, _f32, _i32, _i64, _i32, _p, _p, _p, _p, _p, _i32, _p, _p, _i32, _p]),
    "nfs_mlp_chain_points_train": (ctypes.c_int, [_p, _f32, _i32, _i64, _i32, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _p,
                                                  _i64, _p, _i32, _p]),
    "nfs_mlp_chain_rays": (ctypes.c_int, [_p, _p, _p, _i64, _i32, _f32, _i32, _i32, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _i64,
                                          _p, _i32, _p]),
    "nfs_render_fused_fwd": (ctypes.c_int, [_p, _p, _p, _i64, _i32, _p, _p, _p, _p, _i32, _p, _i64, _i32, _p, _p, _p, _p, _p, _p,
                                            _p, _p, _p, _p, _p, _p]),
    "nfs_render_fused_fwd_train": (ctypes.c_int, [_p, _p, _p, _p, _i64, _i32, _p, _p, _p, _p, _i32, _p, _i64, _i32, _p, _f32,
                                                  _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "nfs_render_fused_bwd": (ctypes.c_int, [_p, _i32, _p, _i64, _i32, _p, _i64, _i64, _p, _p]),
    "nfs_posenc_bf16": (ctypes.c_int, [_p, _p, _p, _p, _p, _i32, _i64, _i32, _i32, _i32, _i32, _i64, _i32, _p, _p]),
    "nfs_gate_bwd_bf16": (ctypes.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _i32, _i32, _p, _p]),
    "nfs_pack_linear_bf16": (ctypes.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p]),
    "nfs_act_grad_bf16": (ctypes.c_int, [_p, _p, _i64, _i32, _i32, _i32, _i64, _p, _p]),
    "nfs_adam_step": (ctypes.c_int, [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _f32, _i32, _p]),
    "nfs_adam_step_dev": (ctypes.c_int, [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _p, _p, _f32, _i32, _p]),
    "nfs_dp_flags_offset": (ctypes.c_uint64, [_i64]),
    "nfs_dp_buffer_bytes": (ctypes.c_uint64, [_i64]),
    "nfs_dp_alloc": (ctypes.c_int, [_i64, _p]),
    "nfs_dp_free": (ctypes.c_int, [_p]),
    "nfs_dp_ipc_export": (ctypes.c_int, [_p, _p]),
    "nfs_dp_ipc_open": (ctypes.c_int, [_p, _p]),
    "nfs_dp_ipc_close": (ctypes.c_int, [_p]),
    "nfs_dp_wait_readers": (ctypes.c_int, [_p, _i32, _i32, _i64, _p, _p]),
    "nfs_dp_adam_step": (ctypes.c_int, [_p, _p, _i32, _i32, _p, _p, _i64, _f32, _f32, _f32, _f32, _p, _p, _f32, _i32, _p, _p,
                                        _p]),
    "nfs_sample_hierarchical": (ctypes.c_int, [_p, _p, _p, _p, _p, _i64, _p, _i64, _i32, _i32, _p, _p, _p, _p, _p, _p]),
}



class WgradJob(ctypes.Structure):
    """struct nfs_wgrad_job of include/nfs_b200.h."""
    _fields_ = [("u_bf16", _p), ("u_pitch", _i64), ("v_bf16", _p), ("v_pitch", _i64), ("n_points", _i64),
                ("m_dim", _i32), ("n_dim", _i32), ("m_valid", _i32), ("n_valid", _i32),
                ("dw", _p), ("ld_m", _i64), ("ld_n", _i64), ("colsum", _p), ("colsum_of_v", _i32)]


class ChainModel(ctypes.Structure):
    """struct nfs_chain_model of include/nfs_b200.h."""
    _fields_ = [("n_layers", _i32), ("k_dims", _p), ("n_dims", _p), ("acts", _p), ("row0", _p), ("w_stack_bf16", _p),
                ("w_rows", _i32), ("bias_terms_bf16", _p), ("freq0", _f32), ("n_octaves", _i32)]


class ChainTrain(ctypes.Structure):
    """struct nfs_chain_train of include/nfs_b200.h."""
    _fields_ = [("x_bf16", _p), ("save_bf16", _p), ("relu_bits", _p), ("rows_per_layer", _i64), ("row0", _i64 * 2)]


class RenderPass(ctypes.Structure):
    """struct nfs_render_pass of include/nfs_b200.h."""
    _fields_ = [("rgb_sigma", _p), ("z_vals", _p), ("g_rgb", _p), ("g_depth", _p), ("n_samples", _i32), ("row0", _i64)]


class ChainBackward(ctypes.Structure):
    """struct nfs_chain_backward of include/nfs_b200.h."""
    _fields_ = [("n_layers", _i32), ("k_dims", _p), ("n_dims", _p), ("acts", _p), ("row0", _p), ("wt_stack_bf16", _p),
                ("w_rows", _i32), ("relu_bits_in", _p), ("bits_rows_per_layer", _i64), ("mask_idx", _p), ("dys_bf16", _p),
                ("save_rows_per_layer", _i64), ("jobs", _p), ("n_jobs", _i32), ("job_waits", _p), ("quad_flags", _p),
                ("producer_pairs", _i32)]


_lib = None


def declared_symbols(header_path=HEADER_PATH):
    """Names of every function the public header declares."""
    with open(header_path) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nfs_[a-z0-9_]+)\s*\(", text)))


def load():
    """dlopen the in-tree library (once) and attach prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "nfs_b200: %s is missing - run `python __graft_entry__.py build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for this path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise RuntimeError("nfs_b200: %s does not export %s (stale build?)" % (LIB_PATH, name))
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().nfs_last_error_string().decode("utf-8", "replace")


def launch_count():
    return int(load().nfs_launch_count())


def call(name, *args):
    """Invoke an int-returning entry point; raise RuntimeError on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (status %d): %s" % (name, rc, last_error()))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
