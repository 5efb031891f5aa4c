"""Turn the ncu reports / launch lists in gpurun_out/ into the small tracked summaries under
profiles/ (round-tagged): per-kernel duration, DRAM bytes, throughput percentages, tensor-pipe
activity, registers, occupancy, top stall reasons."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEYS = [
    ("duration_us", "gpu__time_duration.sum"),
    ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct_of_peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct_of_peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pipe_active_pct", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("tmem_inst_pct", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active"),
    # metrics that DO see tcgen05 cta_group::2 MMAs (the pct above under-reports them):
    ("tensor_hmma_subpipe_cycles", "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"),
    ("tensor_ops_bf16_pct_of_peak", "sm__ops_path_tensor_src_bf16_dst_fp32.avg.pct_of_peak_sustained_elapsed"),
    ("utcmma_inst", "smsp__sass_inst_executed_op_utcmma.sum"),
    ("sm_cycles_elapsed", "sm__cycles_elapsed.avg"),
    ("l1_to_l2_write_MB", "l1tex__m_l1tex2xbar_write_bytes.sum"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("regs", "launch__registers_per_thread"), ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("sm_clock_GHz", "sm__cycles_elapsed.avg.per_second"),
    ("inst_executed", "smsp__inst_executed.sum"),
    ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
]


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def unit_scale(unit, want):
    table = {("ms", "us"): 1e3, ("us", "us"): 1.0, ("ns", "us"): 1e-3, ("s", "us"): 1e6,
             ("Gbyte", "MB"): 1e3, ("Mbyte", "MB"): 1.0, ("Kbyte", "MB"): 1e-3, ("byte", "MB"): 1e-6,
             ("Kbyte", "KB"): 1.0, ("byte", "KB"): 1e-3, ("Mbyte", "KB"): 1e3,
             ("Ghz", "GHz"): 1.0, ("Mhz", "GHz"): 1e-3, ("hz", "GHz"): 1e-9, ("GHz", "GHz"): 1.0}
    return table.get((unit, want))


def summarize(rep, name):
    hdr, units, rows = raw_rows(rep)
    out = []
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in rows:
        kn = r[hdr.index("Kernel Name")]
        short = kn.split("(")[0].split("::")[-1]
        rec = {"kernel": short, "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for label, key in KEYS:
            if key not in hdr:
                continue
            i = hdr.index(key)
            try:
                v = float(r[i])
            except ValueError:
                continue
            want = label.rsplit("_", 1)[-1]
            sc = unit_scale(units[i], want)
            rec[label] = round(v * sc, 3) if sc else round(v, 3)
        stalls = []
        for i in stall_cols:
            try:
                stalls.append((float(r[i]), hdr[i].split("stalled_")[1].split("_per_issue")[0]))
            except ValueError:
                pass
        if "tensor_hmma_subpipe_cycles" in rec and rec.get("sm_cycles_elapsed"):
            # the sub-pipe counter is summed over the SM's four sub-partitions
            rec["tensor_active_frac"] = round(rec["tensor_hmma_subpipe_cycles"] / 4.0 / rec["sm_cycles_elapsed"], 3)
        rec["top_stalls_per_issue"] = [[n, round(v, 2)] for v, n in sorted(stalls, reverse=True)[:4]]
        out.append(rec)
    path = os.path.join(OUT, "%s_%s_ncu_full_summary.json" % (TAG, name))
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, len(out), "kernels")
    return out


def launches(csv_path, name):
    """ncu --metrics gpu__time_duration.sum launch list -> per-kernel totals and shares."""
    lines = [l for l in open(csv_path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}.get(r[mu], 1.0)
        k = r[kn].split("(")[0].split("::")[-1]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    path = os.path.join(OUT, "%s_%s_launches_summary.csv" % (TAG, name))
    with open(path, "w") as f:
        f.write("kernel,launches,total_us,share_pct\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.1f\n" % (k, n, t, 100 * t / tot))
    print("wrote", path)


if __name__ == "__main__" and len(sys.argv) > 3:
    # python scripts/summarize_ncu.py <tag> <report.ncu-rep> <name>
    summarize(sys.argv[2], sys.argv[3])
elif __name__ == "__main__":
    g = os.path.join(ROOT, "gpurun_out")
    for name in ("k1", "k3"):
        rep = os.path.join(g, "prof_%s.ncu-rep" % name)
        if os.path.exists(rep):
            summarize(rep, name)
        lst = os.path.join(g, "launches_%s.csv" % name)
        if os.path.exists(lst):
            launches(lst, name)
