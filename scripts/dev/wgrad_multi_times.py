"""Per-job start / end times of the multi-job weight-gradient launch of a cfg 4 (NeRFWithDINO) step (NFS_WGRAD_TIMES):
which jobs finish last - the input for the cost weights in nfs_wgrad_multi_bf16."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.argv = [sys.argv[0]]
    os.environ["NFS_WGRAD_TIMES"] = "1"
    src = os.path.join(ROOT, "scripts", "profile_step_g3.py")
    exec(compile(open(src).read().split("reps = 5")[0], src, "exec"), dict(globals(), __file__=src))
    sys.exit(0)
out = subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True)
lines = [l for l in out.stdout.splitlines() if l.startswith("wgtimes")]
jobs = len(set(l.split()[2] for l in lines))
rows = [l.split() for l in lines[-jobs:]]               # the last step's launch
if not rows:
    print(out.stdout[-1500:], out.stderr[-1500:]); sys.exit(1)
t0 = min(int(r[r.index("start") + 1]) for r in rows)
for r in sorted(rows, key=lambda r: int(r[2])):
    g = lambda k: r[r.index(k) + 1]
    print("job %2s paired %s ctas %3s  M %3s N %3s P %6s  start %6.1f us  end %6.1f us" % (
        g("job"), g("paired"), g("ctas"), g("M"), g("N"), g("P"), (int(g("start")) - t0) / 1e3, (int(g("end")) - t0) / 1e3))
