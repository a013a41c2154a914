"""Turn the ncu captures in gpurun_out/ into the small, tracked summaries under profiles/.

    python profiles/summarize.py r1b      # reads gpurun_out/prof_*_r1b.ncu-rep and launches_r1b.csv

Writes profiles/<tag>_<kernel>.json (selected metrics of one `ncu --set full` launch),
profiles/<tag>_launches.txt (per-kernel time shares of one bench step under
`--metrics gpu__time_duration.sum`) and profiles/ncu_traffic.json (DRAM bytes per launch of the
kernels bench.py reports a roofline for). The .ncu-rep files themselves stay in gpurun_out/ (scratch).
"""
import csv
import glob
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__inst_executed_op_shared_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "lts__t_sector_op_red_hit_rate.pct", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum",
    "sm__pipe_tensor_op_umma_cycles_active.avg.pct_of_peak_sustained_elapsed",
]


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {n: (vals[i], units[i]) for i, n in enumerate(hdr)}


def main(tag):
    traffic = {}
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_*_{tag}.ncu-rep"))):
        kern = os.path.basename(rep)[len("prof_"):-len(f"_{tag}.ncu-rep")]
        m = raw_metrics(rep)
        summ = {"kernel": m.get("Kernel Name", ("?",))[0], "capture": f"ncu --set full --clock-control none, one launch, {tag}"}
        for k in KEEP:
            if k in m:
                summ[k] = {"value": m[k][0], "unit": m[k][1]}
        with open(os.path.join(OUT, f"{tag}_{kern}.json"), "w") as f:
            json.dump(summ, f, indent=1)

        def to_bytes(key):
            v, u = m[key]
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        traffic[kern] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        print(kern, summ.get("gpu__time_duration.sum"), "dram bytes", traffic[kern])
    if traffic:
        # merge: a partial re-capture (a few kernels under a new tag) must not drop the other kernels' entries
        tp = os.path.join(OUT, "ncu_traffic.json")
        old = json.load(open(tp)) if os.path.exists(tp) else {}
        tags = old.get("tags", {k: old.get("tag") for k in old if k not in ("tag", "tags")})
        tags.update({k: tag for k in traffic})
        merged = {k: v for k, v in old.items() if k not in ("tag", "tags")}
        merged.update(traffic)
        with open(tp, "w") as f:
            json.dump({"tag": tags.get("fac_bwd_march", tag), "tags": tags, **merged}, f, indent=1)
    lf = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if os.path.exists(lf):
        rows = list(csv.reader(l for l in open(lf) if l.startswith('"')))
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = defaultdict(list)
        for r in rows[1:]:
            agg[r[ki]].append(float(r[vi].replace(",", "")))
        tot = sum(sum(v) for v in agg.values())
        with open(os.path.join(OUT, f"{tag}_launches.txt"), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 3 --warmup 3 --kernels-only\n")
            f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
            f.write(f"{'total_us':>10} {'n':>4} {'avg_us':>9} {'share':>7}  kernel\n")
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
                f.write(f"{sum(v) / 1e3:10.1f} {len(v):4d} {sum(v) / len(v) / 1e3:9.1f} {100 * sum(v) / tot:6.1f}%  {k[:110]}\n")
        print(open(os.path.join(OUT, f"{tag}_launches.txt")).read())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r1b")
