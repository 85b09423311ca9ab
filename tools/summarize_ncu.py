"""Turns ncu exports (run under gpurun) into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r01_launches.md
  python tools/summarize_ncu.py kernel   gpurun_out/prof_conv_r1.ncu-rep profiles/r01_conv_ncu.md
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sm__sass_l1tex_t_requests_pipe_lsu_mem_global_op_ldgsts.sum", "smsp__inst_executed.sum"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, data = None, []
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for x in data:
        a = agg.setdefault(x["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += float(x["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write(f"source: {src}; {len(data)} launches, {tot / 1e3:.1f} us total\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:110]}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.2f} | {100 * v[1] / tot:.1f}% |\n")


def kernel(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in rows[2:]:
            f.write(f"## {r[idx['Kernel Name']][:120]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
