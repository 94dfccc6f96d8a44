"""Turns an `ncu --set full` capture of the covariance kernels (taken on the GPU box, read here)
into profiles/ncu_<tag>_summary.csv and profiles/ncu_traffic.json - the per-step DRAM traffic
bench.py reports as roofline.traffic / roofline_staging.traffic.

    ncu -i gpurun_out/prof_r02.ncu-rep --page raw --csv > /tmp/prof.csv
    python scripts/ncu_traffic.py /tmp/prof.csv r02 "ncu --set full ... python bench.py --steps 2 ..."
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag = sys.argv[1], sys.argv[2]
cmd = sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = [hdr.index(w) for w in want if w in hdr]
out = os.path.join(ROOT, "profiles", "ncu_%s_summary.csv" % tag)
recs = []
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
        recs.append({hdr[i]: r[i] for i in idx})


def gb(r):
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    ur = units[hdr.index("dram__bytes_read.sum")]
    uw = units[hdr.index("dram__bytes_write.sum")]
    return float(r["dram__bytes_read.sum"]) * scale[ur] + float(r["dram__bytes_write.sum"]) * scale[uw]


# one covariance pass of the benchmark = the two full-size staging launches + the two full-size
# contraction launches (the largest instance of each kernel in the capture)
def biggest(pred, n):
    c = sorted((r for r in recs if pred(r["Kernel Name"])), key=gb, reverse=True)
    return c[:n]


# phase 0 = the TMA-fed kernel (stage_tma_kernel) when the capture has it, else the largest
# stage_group_kernel launch; phase 1 = the largest stage_group_kernel launch that is not phase 0
tma = biggest(lambda k: "stage_tma_kernel" in k, 1)
stage = biggest(lambda k: "stage_group_kernel" in k, 4)
if tma:
    ph0 = tma[0]
    ph1 = stage[0] if stage else None
else:
    ph0 = stage[0]
    ph1 = next((r for r in stage if gb(r) < 0.5 * gb(ph0)), None)
gen = biggest(lambda k: "contraction_tc_kernel" in k, 1)[0]
ac = biggest(lambda k: "autocorr_tc_kernel" in k, 1)[0]
traffic = {
    "source": "profiles/ncu_%s_summary.csv (%s)" % (tag, cmd),
    "staging_bytes_per_step": gb(ph0) + (gb(ph1) if ph1 else 0.0),
    "contraction_bytes_per_step": gb(gen) + gb(ac),
    "per_launch": {"stage_phase0": gb(ph0), "stage_phase1": gb(ph1) if ph1 else None,
                   "contraction_generic": gb(gen), "contraction_sliding_window": gb(ac)},
    "tensor_pipe_pct": {
        "contraction_generic": float(gen["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]),
        "contraction_sliding_window": float(ac["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"])},
}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
print("wrote", out)
