"""Turns ncu exports (run under gpurun) into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r01_launches.md
  python tools/summarize_ncu.py kernel   gpurun_out/prof_step_r1_raw.csv profiles/r01_step_ncu.md
  python tools/summarize_ncu.py traffic  gpurun_out/prof_step_r1_raw.csv profiles/r01_traffic.json
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
        "sm__sass_l1tex_t_requests_pipe_lsu_mem_global_op_ldgsts.sum", "smsp__inst_executed.sum",
        # shared-memory side (round 2): wavefronts of 128 B read by the tensor cores / moved by LSU (LDGSTS, LDS, STS)
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum",
        "smsp__sass_inst_executed_op_utcmma.sum", "launch__occupancy_limit_shared_mem"]


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
    """src: a .ncu-rep, or the CSV `ncu -i rep --page raw --csv` wrote on the GPU box (reports with more than a
    few kernels exceed what gpurun copies back)."""
    if src.endswith(".csv"):
        out = open(src).read()
    else:
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


def traffic(src, dst):
    """DRAM bytes per launch of the conv kernels (roofline.traffic of bench.py), keyed by "CINxCOUT"."""
    import json
    import re
    rows = list(csv.reader(open(src).read().splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    kern = {}
    for r in rows[2:]:
        m = re.search(r"conv_fwd_tc<\(int\)(\d+), \(int\)(\d+)", r[idx["Kernel Name"]]) or \
            re.search(r"conv_fwd_tc<(\d+), (\d+)", r[idx["Kernel Name"]])
        if not m:
            continue
        def val(name):
            v, u = float(r[idx[name]]), rows[1][idx[name]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        e = kern.setdefault(f"{m.group(1)}x{m.group(2)}", {"launches": []})
        e["launches"].append({"dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                              "us": float(r[idx["gpu__time_duration.sum"]]), "grid": int(r[idx["launch__grid_size"]])})
    for e in kern.values():
        e["dram_bytes_per_launch"] = max(l["dram_bytes"] for l in e["launches"])
    json.dump({"source": "ncu --set full --clock-control none (" + src + "); one un-graphed step of bench.py, the kernels run "
               "back to back so the L2 is warm from the previous layer; dram_bytes_per_launch = the largest launch of the shape",
               "kernels": kern}, open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
