#!/usr/bin/env python
"""Condense ncu outputs from gpurun_out/ into small, committed summaries under profiles/.

  summarize_ncu.py launches <launches.csv> <out.md>        per-kernel launch count / time / share
  summarize_ncu.py full <prof.ncu-rep> <out.md>            key metrics of each profiled launch
  summarize_ncu.py traffic <dram.csv> <out.json> <prec>    average DRAM bytes per launch per kernel family
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_active.avg",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name[:70]


def launches(path, out):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
        n += 1
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n{n} launches captured, {total / 1e6:.3f} ms of device time "
                "(ncu-serialised, cold cache: compare SHARES, not absolutes)\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / c / 1e3:.1f} |\n")
    print(open(out).read())


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n\n")
        for d in data:
            rec = dict(zip(hdr, d))
            f.write(f"## launch {rec['ID']}: `{short(rec['Kernel Name'])}` grid {rec['Grid Size']} block {rec['Block Size']}\n\n")
            f.write("| metric | value | unit |\n|---|---:|---|\n")
            for i, h in enumerate(hdr):
                base = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h
                if any(h == k or h.endswith(k) for k in KEYS):
                    f.write(f"| {h} | {d[i]} | {units[i]} |\n")
            f.write("\n")
    print(open(out).read()[:6000])


def traffic(path, out_json, precision):
    """ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv log -> per-kernel-family
    average DRAM bytes per launch, merged into profiles/kernel_traffic.json under `precision`."""
    import json, os
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = {}
    for r in rd:
        name = r["Kernel Name"]
        for form in ("conv_stream_tma_pair_kernel", "conv_stream_tma_kernel", "conv_stream_pair_kernel"):
            name = name.replace(form, "conv_stream_kernel")   # one family, four forms (bench.py's label)
        name = name.replace("lstm_pair_kernel", "lstm_tc_kernel")
        fam = next((f for f in ("conv_stream_kernel", "ru_persist_kernel", "ru_group_kernel", "ru_pair_kernel", "conv1d_tc_kernel",
                                "conv1d_f32_kernel", "lstm_tc_kernel", "stem_conv", "tail_conv", "vq_scan_kernel", "vq_encode_kernel", "snake_aa2_kernel",
                                "snake_kernel") if f in name), None)
        if fam is None:
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}.get(unit, 1.0)
        a = agg.setdefault(fam, {"launch_ids": set(), "rd": 0.0, "wr": 0.0, "ns": 0.0})
        a["launch_ids"].add(r["ID"])
        if r["Metric Name"] == "dram__bytes_read.sum":
            a["rd"] += v * scale
        elif r["Metric Name"] == "dram__bytes_write.sum":
            a["wr"] += v * scale
        elif r["Metric Name"] == "gpu__time_duration.sum":
            a["ns"] += v * scale
    res = {}
    for fam, a in agg.items():
        n = len(a["launch_ids"])
        res[fam] = {"launches": n, "dram_bytes_per_launch": (a["rd"] + a["wr"]) / n, "dram_read_per_launch": a["rd"] / n,
                    "dram_write_per_launch": a["wr"] / n, "ncu_us_per_launch": a["ns"] / n / 1e3}
    cur = {}
    if os.path.exists(out_json):
        cur = json.load(open(out_json))
    cur[precision] = res
    cur["_how"] = ("ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over "
                   "`BC_LSTM_WAVEFRONT=0 bench.py --clips-per-gpu 8 --steps 1` (micro-batch 8: the launch shapes of the full bench's conv kernels; "
                   "the small-batch LSTM wave front is switched off so that the LSTM and its input projection are single launches as in the full bench)")
    json.dump(cur, open(out_json, "w"), indent=1, sort_keys=True)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
