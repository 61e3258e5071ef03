"""ncu report -> the per-kernel summary bench.py and DESIGN.md quote (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/<rep>.ncu-rep <workload> [--merge profiles/ncu_traffic.json] [--details profiles/<name>.csv]
Per kernel (the LAST launch of every kernel name in the report): DRAM bytes (read + write), duration, issue-active
percentage, warp instructions, achieved warps per SM, registers."""
import csv
import json
import subprocess
import sys

KEYS = {"project_kernel": "project", "project3d_frames_kernel": "project", "depth_rank_bucket_kernel": "depth_rank",
        "depth_rank_kernel": "depth_rank_radix_fallback", "partition_kernel": "partition",
        "sort_lists_kernel": "sort_lists", "sort_split_kernel": "sort_split", "block_lists_kernel": "block_lists",
        "raster_fwd6_kernel": "raster_fwd", "raster_fwd_kernel": "raster_fwd", "raster_bwd3_kernel": "raster_bwd",
        "raster_bwd2_kernel": "raster_bwd_v5", "project_bwd_kernel": "project_bwd", "fill_empty_kernel": "fill_empty",
        "ssim_fwd_kernel": "ssim_fwd", "loss_bwd_kernel": "loss_bwd", "build_worklist_kernel": "build_worklist"}
METRICS = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
           "smsp__issue_active.avg.pct": "issue_active_pct", "smsp__issue_active.avg.per_cycle_active": "issue_per_cycle",
           "smsp__inst_executed.sum": "inst_executed", "sm__warps_active.avg.per_cycle_active": "achieved_warps_per_sm",
           "launch__registers_per_thread": "registers", "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
           "launch__grid_size": "grid"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0,
        "second": 1e3}


def main():
    rep, wl = sys.argv[1], sys.argv[2]
    merge = sys.argv[sys.argv.index("--merge") + 1] if "--merge" in sys.argv else None
    details = sys.argv[sys.argv.index("--details") + 1] if "--details" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = {}
    keep_cols = [ix["Kernel Name"]] + [ix[m] for m in METRICS if m in ix]
    kept = [[hdr[i] for i in keep_cols], [units[i] for i in keep_cols]]
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        key = next((v for k, v in KEYS.items() if k + "<" in name or k + "(" in name), None)
        if key is None:
            continue
        e = {}
        for m, short in METRICS.items():
            if m not in ix:
                continue
            try:
                val = float(r[ix[m]].replace(",", ""))
            except ValueError:
                continue
            e[short] = val * UNIT.get(units[ix[m]], 1.0) if short in ("duration", "dram_read", "dram_write") else val
        out[key] = {"dram_bytes_per_launch": int(e.get("dram_read", 0) + e.get("dram_write", 0)), "duration_ms": e.get("duration"),
                    "issue_active_pct": e.get("issue_active_pct", 100.0 * e.get("issue_per_cycle", 0.0)),
                    "inst_executed": int(e.get("inst_executed", 0)), "achieved_warps_per_sm": e.get("achieved_warps_per_sm"),
                    "registers": e.get("registers"), "threads_per_inst": e.get("threads_per_inst"), "cuda_kernel": name.split("(")[0][-60:]}
        kept.append([r[i] for i in keep_cols])
    print(json.dumps(out, indent=1))
    if details:
        with open(details, "w", newline="") as f:
            csv.writer(f).writerows(kept)
    if merge:
        try:
            cur = json.load(open(merge))
        except Exception:
            cur = {}
        cur.setdefault(wl, {}).update(out)
        cur["_source_" + wl] = f"ncu --set full --clock-control none, {rep}; last launch of every kernel; tools/ncu_summary.py"
        json.dump(cur, open(merge, "w"), indent=1)


if __name__ == "__main__":
    main()
