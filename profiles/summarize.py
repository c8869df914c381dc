"""Turn gpurun_out/*.ncu-rep and launch-list CSVs into the small text summaries committed here.

  python profiles/summarize.py full  gpurun_out/r01_nerf_train_kernels.ncu-rep  profiles/r01_nerf_train_kernels_ncu_full.txt
  python profiles/summarize.py list  gpurun_out/r01_launches_nerf_train_bf16.csv profiles/r01_launches_nerf_train_bf16.txt
"""
import csv
import subprocess
import sys
from collections import OrderedDict

WANT = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; source report {rep}\n")
        f.write("# one block per captured launch; values are per launch (cold caches, serialised)\n")
        for r in rows[2:]:
            f.write(f"\n== {r[ik][:90]}\n")
            for name in WANT:
                if name in hdr:
                    i = hdr.index(name)
                    f.write(f"   {name:90s} {r[i]:>16s} {units[i]}\n")


def launch_list(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ours = [(r[ik], float(r[iv].replace(",", "")) / 1e3) for r in rows[1:] if not r[ik].startswith("void at")
            and "at_cuda_detail" not in r[ik]]
    # last complete step: from the last-but-one sample_coarse_kernel to the last one
    idx = [i for i, (n, _) in enumerate(ours) if "sample_coarse_kernel" in n]
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; source {path}\n")
        f.write("# per-launch device time in us (cold-cache, serialised: compare SHARES, not absolutes)\n")
        if len(idx) >= 2:
            a, b = idx[-2], idx[-1]
            step = ours[a:b]
            tot = sum(v for _, v in step)
            f.write(f"# one full step = {len(step)} launches of our kernels, {tot:.1f} us in total\n\n")
            agg = OrderedDict()
            for n, v in step:
                key = n.split("(")[0].replace("void ", "").replace("lnrf::", "")
                agg.setdefault(key, [0, 0.0])
                agg[key][0] += 1
                agg[key][1] += v
            f.write(f"{'kernel':48s} {'launches':>8s} {'us':>12s} {'share':>8s}\n")
            for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{k[:48]:48s} {c:8d} {v:12.1f} {100 * v / tot:7.1f}%\n")
            f.write("\n# in launch order\n")
            for n, v in step:
                f.write(f"{v:12.1f}  {n[:100]}\n")
        else:
            for n, v in ours:
                f.write(f"{v:12.1f}  {n[:100]}\n")


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
