"""ncu launch list of the headline step (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch,
scripts/profile_r02*.sh) -> profiles/<tag>_step_launches_summary.json + profiles/traffic_step.json: the launches of ONE
eager training step (from one pack_stack_kernel to the next), per-kernel device time, share and DRAM bytes.
usage: python scripts/summarize_step.py gpurun_out/r2b_launches_step.csv r02b"""
import csv, json, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag = sys.argv[1], sys.argv[2]
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
hdr = rows[0]
I = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
launch = collections.OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(int(r[I["ID"]]), {"name": r[I["Kernel Name"]]})
    v = float(r[I["Metric Value"]].replace(",", ""))
    u = r[I["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[r[I["Metric Name"]]] = v
ids = list(launch)
starts = [i for i in ids if "pack_stack_kernel" in launch[i]["name"]]
lo, hi = (starts[1], starts[2]) if len(starts) >= 3 else (starts[-1], ids[-1] + 1)
step = [launch[i] for i in ids if lo <= i < hi]


def short(n):
    n = n.split("(")[0].replace("void ", "")
    for p in ("nfs::<unnamed>::", "at::native::", "at::"):
        n = n.replace(p, "")
    return n[:80]


agg = collections.OrderedDict()
for d in step:
    a = agg.setdefault(short(d["name"]), [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
traffic = sum(a[2] + a[3] for a in agg.values())
out = {"command": "python bench.py --quick --no-extras --no-cpu-baseline --steps 2 --warmup 3 (one eager warm-up step of "
                  "GraphedTrainStep; ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none)",
       "launches_per_step": len(step), "kernel_time_us_serialised": round(tot, 1), "dram_bytes_per_step": traffic,
       "dram_KB_per_point": round(traffic / 1048576 / 1e3, 2),
       "kernels": [{"kernel": k, "launches": a[0], "us": round(a[1], 1), "share_pct": round(100 * a[1] / tot, 2),
                    "dram_read_MB": round(a[2] / 1e6, 1), "dram_write_MB": round(a[3] / 1e6, 1)}
                   for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
path = os.path.join(ROOT, "profiles", "%s_step_launches_summary.json" % tag)
json.dump(out, open(path, "w"), indent=1)
json.dump({"dram_bytes_per_step": traffic, "source": "profiles/%s_step_launches_summary.json (ncu dram__bytes_read.sum + "
           "dram__bytes_write.sum over all launches of one training step)" % tag},
          open(os.path.join(ROOT, "profiles", "traffic_step.json"), "w"), indent=1)
print("wrote", path, "launches", len(step), "kernel time %.1f us" % tot, "DRAM %.2f GB = %.2f KB/point" % (traffic / 1e9, traffic / 1048576 / 1e3))
for k in out["kernels"][:8]:
    print(k)
