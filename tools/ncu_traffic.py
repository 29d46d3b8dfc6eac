"""Summaries of the round's ncu captures (tools/profile_round.sh).

    python tools/ncu_traffic.py gpurun_out profiles

* `<src>/r02_launches_bench.csv` -- the single-pass list of every launch of a default bench.py
  run (duration + DRAM bytes per launch, the kernels at the bench's own sizes) ->
  `profiles/r02_launches_bench_summary.json` (per kernel: launches, total / largest duration,
  DRAM bytes of the largest launches) and `profiles/r02_dram_traffic.json`, which bench.py
  reads for the `traffic` field of its rooflines.
* `<src>/r02_<name>_raw.csv` -- the raw page of one `--set full` capture per dominant kernel ->
  `profiles/r02_<name>_ncu_summary.json` (duration, DRAM bytes, instructions, IPC, occupancy,
  top stall reasons).
"""
import collections
import csv
import json
import os
import statistics
import sys


def number(entry):
    if entry is None:
        return None
    unit, value = entry
    try:
        x = float(value.replace(",", ""))
    except ValueError:
        return None
    factor = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12, "us": 1e-6, "ms": 1e-3,
              "ns": 1e-9, "s": 1}.get(unit.split("/")[0], 1)
    return x * factor


def summarise_raw(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 20]
    if len(rows) < 3:
        return None
    d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
    out = {
        "kernel": d.get("Kernel Name", ("", ""))[1],
        "grid": d.get("Grid Size", ("", ""))[1], "block": d.get("Block Size", ("", ""))[1],
        "duration_s": number(d.get("gpu__time_duration.sum")),
        "dram_bytes_read": number(d.get("dram__bytes_read.sum")),
        "dram_bytes_write": number(d.get("dram__bytes_write.sum")),
        "warp_instructions": number(d.get("smsp__inst_executed.sum")),
        "ipc_active": number(d.get("sm__inst_executed.avg.per_cycle_active")),
        "issue_slots_busy_pct": number(d.get("sm__inst_issued.avg.pct_of_peak_sustained_active")),
        "achieved_occupancy_pct": number(d.get("sm__warps_active.avg.pct_of_peak_sustained_active")),
        "registers_per_thread": number(d.get("launch__registers_per_thread")),
        "dram_throughput_pct": number(d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
        "l1tex_data_pipe_lsu_wavefronts_mem_shared_pct": number(
            d.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed")),
        "fp64_pipe_pct": number(d.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")),
    }
    stalls = {}
    for k, (u, v) in d.items():
        if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
            try:
                stalls[k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = round(float(v), 2)
            except ValueError:
                pass
    out["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
    if out["dram_bytes_read"] is not None and out["dram_bytes_write"] is not None:
        out["dram_bytes_per_launch"] = out["dram_bytes_read"] + out["dram_bytes_write"]
    return out


def read_launches(path):
    """[{name, ns, bytes}] in launch order."""
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    head = rows[0]
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = dict(zip(head, r))
        e = launches.setdefault(int(d["ID"]), {"name": d["Kernel Name"], "grid": d["Grid Size"], "ns": 0.0, "bytes": 0.0})
        v = float(d["Metric Value"].replace(",", ""))
        if d["Metric Name"] == "gpu__time_duration.sum":
            e["ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(d["Metric Unit"], 1)
        else:
            e["bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(d["Metric Unit"], 1)
    return [launches[k] for k in sorted(launches)]


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0]


def largest(entries, key="bytes"):
    """Median of the entries within 20 % of the largest (the launches at the bench's full size)."""
    top = max(e[key] for e in entries)
    return [e for e in entries if e[key] >= 0.8 * top]


# roofline key of bench.py -> kernel-name fragment (one launch per call)
PER_LAUNCH = {
    "gm_fused_cfg2_16384": "gm_fused",
    "temporal_sum": "temporal_stream_kernel<float, float, float, 0>",
    "temporal_max": "temporal_stream_kernel<float, float, float, 3>",
    "temporal_mean": "temporal_stream_kernel<float, float, float, 4>",
    "temporal_sum_int16": "temporal_stream_kernel<short, double, int, 0>",
    "temporal_max_int16": "temporal_stream_kernel<short, float, short, 3>",
    "temporal_std": "temporal_moments_stream_kernel<float",
    "temporal_median": "temporal_sort_reg_kernel<float",
    "cumulative_sum": "temporal_cumulative_stream_kernel<float",
    "smooth_fast_kernel": "smooth_fast_kernel<float",
    "moving_max_block_kernel": "moving_max_block_kernel<float",
    "hillshade_quad_kernel": "hillshade_quad_kernel<float",
}
# zonal statistics: one CALL = the launches from one call's first kernel to the next call's
PER_CALL = {
    "zonal_mean": "zonal_reduce_warp_kernel<float, 1>",
    "zonal_max": "zonal_reduce_warp_kernel<float, 4>",
    "zonal_p90": "zonal_select_main_kernel<float>",
}


def main(src, dst):
    traffic = {}
    path = os.path.join(src, "r02_launches_bench.csv")
    if os.path.exists(path):
        launches = read_launches(path)
        ours = [e for e in launches if "gm::" in e["name"] or e["name"].startswith("gm_")]
        by = collections.defaultdict(list)
        for e in ours:
            by[short(e["name"])].append(e)
        total = sum(e["ns"] for e in ours)
        summary = {"command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                              "--clock-control none python bench.py --steps 5 --warmup 3 --leg-steps 3 --e2e-steps 2",
                   "note": "single pass, no replay; durations are serialised launches (shares, not absolutes)",
                   "kernels": []}
        for name, es in sorted(by.items(), key=lambda kv: -sum(e["ns"] for e in kv[1])):
            big = largest(es, "ns")
            summary["kernels"].append({
                "kernel": name, "launches": len(es), "total_ms": round(sum(e["ns"] for e in es) / 1e6, 3),
                "share_of_our_kernel_time": round(sum(e["ns"] for e in es) / total, 4),
                "largest_launch_ms": round(statistics.median(e["ns"] for e in big) / 1e6, 4),
                "largest_launch_dram_bytes": statistics.median(e["bytes"] for e in big)})
        json.dump(summary, open(os.path.join(dst, "r02_launches_bench_summary.json"), "w"), indent=1)
        source = "profiles/r02_launches_bench.csv (single-pass ncu of bench.py)"
        for key, fragment in PER_LAUNCH.items():
            es = [e for e in ours if fragment in e["name"]]
            if not es:
                continue
            big = largest(es)
            traffic[key] = {"dram_bytes_per_launch": statistics.median(e["bytes"] for e in big),
                            "kernel_duration_s_under_ncu": statistics.median(e["ns"] for e in big) / 1e9,
                            "launches_counted": len(big), "source": source}
        # a call starts with its preparation (first call on a soup and grid) or, when a resident
        # soup kept that, with the first kernel of the statistic itself
        FIRST = ("zonal_reduce_warp_kernel", "zonal_select_bracket_kernel")
        calls, current, opened_by_prepare = [], None, False
        for e in ours:
            starts = "poly_transform_kernel" in e["name"]
            if any(f in e["name"] for f in FIRST):
                starts = not (opened_by_prepare and current is not None
                              and not any(any(f in x["name"] for f in FIRST) for x in current))
            if starts:
                current = []
                calls.append(current)
                opened_by_prepare = "poly_transform_kernel" in e["name"]
            if current is not None and ("zonal" in e["name"] or "poly_" in e["name"]):
                current.append(e)
        for key, fragment in PER_CALL.items():
            mine = [c for c in calls if any(fragment in e["name"] for e in c)]
            if not mine:
                continue
            sums = [{"bytes": sum(e["bytes"] for e in c), "ns": sum(e["ns"] for e in c)} for c in mine]
            big = largest(sums)
            traffic[key] = {"dram_bytes_per_launch": statistics.median(s["bytes"] for s in big),
                            "kernel_duration_s_under_ncu": statistics.median(s["ns"] for s in big) / 1e9,
                            "launches_counted": len(big), "per": "aggregate_polygons call (all its kernels)",
                            "source": source}
    for name in sorted(os.listdir(src)):
        if not (name.startswith("r02_") and name.endswith("_raw.csv")):
            continue
        summary = summarise_raw(os.path.join(src, name))
        if summary is None:
            continue
        stem = name[:-len("_raw.csv")]
        json.dump(summary, open(os.path.join(dst, stem + "_ncu_summary.json"), "w"), indent=1)
        print(stem, summary["kernel"][:50], summary["duration_s"], summary.get("dram_bytes_per_launch"))
    json.dump(traffic, open(os.path.join(dst, "r02_dram_traffic.json"), "w"), indent=1)
    for k, v in traffic.items():
        print(k, round(v["dram_bytes_per_launch"] / 1e6, 1), "MB", round(v["kernel_duration_s_under_ncu"] * 1e3, 3), "ms")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
