#!/usr/bin/env python3
"""Turn the round's ncu captures (gpurun_out/, scratch) into the tracked evidence under profiles/:

  python tools/ncu_summary.py r02a            reads  gpurun_out/r02a_counts_<N>.csv   (ncu --metrics ... --csv --log-file)
                                              and    gpurun_out/r02a_step_<N>.ncu-rep (ncu --set full), N in 4096, 65536
                                              writes profiles/r02a_step_kernel_<N>_ncu_raw.txt and profiles/kernel_counts.json

kernel_counts.json carries the per-env-step instruction and flop counts of the walk-ik step kernel that bench.py turns into the
issue-rate and fp32 roofline fractions, plus the measured DRAM bytes per launch.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
GO = os.path.join(ROOT, "gpurun_out")
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def to_num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def read_counts(path):
    """ncu --csv --log-file with --metrics: long format (one row per launch x metric) -> per-launch dicts."""
    if not os.path.exists(path):
        return []
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    launches = {}
    for r in rows:
        d = launches.setdefault(r["ID"], {"kernel": r["Kernel Name"], "grid": r.get("Grid Size")})
        v = to_num(r["Metric Value"])
        unit = r.get("Metric Unit", "")
        if v is not None and unit in ("Kbyte", "KB"):
            v *= 1e3
        if v is not None and unit in ("Mbyte", "MB"):
            v *= 1e6
        if v is not None and unit == "ms":
            v *= 1e3
        if v is not None and unit == "ns":
            v *= 1e-3
        d[r["Metric Name"]] = v
    return list(launches.values())


def raw_page(rep):
    if not os.path.exists(rep):
        return None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (units[i], vals[i]) for i, h in enumerate(hdr)}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02a"
    counts = {}
    for n in (4096, 65536):
        L = [l for l in read_counts(os.path.join(GO, f"{tag}_counts_{n}.csv")) if "step_kernel" in l["kernel"]]
        page = raw_page(os.path.join(GO, f"{tag}_step_{n}.ncu-rep"))
        path = os.path.join(ROOT, "profiles", f"{tag}_step_kernel_{n}_ncu_raw.txt")
        with open(path, "w") as f:
            f.write(f"# step_kernel at {n} envs, walk-ik flat, de-synchronised batch (bench.py --steps 4 --warmup 3 --no-extras --envs-per-gpu {n}, launch 300+).\n")
            if L:
                f.write(f"# (a) ncu --metrics pass over {len(L)} consecutive launches (cold cache, serialised): per-launch means\n")
                keys = [k for k in L[0] if k not in ("kernel", "grid")]
                mean = {k: sum(l.get(k) or 0.0 for l in L) / len(L) for k in keys}
                f.write(f"kernel: {L[0]['kernel']}\n")
                for k in keys:
                    f.write(f"{k:75s} {mean[k]:.6g}\n")
                flop = 2 * mean.get("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", 0) + mean.get("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", 0) \
                    + mean.get("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", 0)
                inst = mean.get("smsp__inst_executed.sum", 0)
                dur = mean.get("gpu__time_duration.sum", 0)
                f.write(f"derived: warp instructions per env-step {inst / n:.1f}; fp32 flop per env-step {flop / n:.0f}; "
                        f"DRAM bytes per launch {mean.get('dram__bytes_read.sum', 0) + mean.get('dram__bytes_write.sum', 0):.0f}; duration {dur:.1f} us\n")
                counts[n] = {"warp_inst_per_env_step": inst / n, "flop_per_env_step": flop / n,
                             "dram_bytes_per_launch": mean.get("dram__bytes_read.sum", 0) + mean.get("dram__bytes_write.sum", 0), "ncu_duration_us": dur,
                             "issue_active_pct": mean.get("smsp__issue_active.avg.pct_of_peak_sustained_active")}
            if page:
                f.write("# (b) ncu --set full --clock-control none, one launch: selected raw metrics, then the warp-state sampling\n")
                f.write(f"kernel: {page.get('Kernel Name', ('', '?'))[1]}\n")
                for k in KEEP:
                    if k in page:
                        f.write(f"{k:75s} {page[k][1]} {page[k][0]}\n")
                st = sorted(((to_num(v[1]) or 0.0, k) for k, v in page.items() if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")), reverse=True)
                tot = sum(x for x, _ in st) or 1.0
                f.write("stall samples (smsp__pcsamp_warps_issue_stalled_*): " + ", ".join(f"{k.split('stalled_')[1]} {100 * x / tot:.1f}%" for x, k in st[:8]) + "\n")
        print("wrote", path)
    if 4096 in counts:
        kc = {"walk_ik_plane": {"warp_inst_per_env_step": round(counts[4096]["warp_inst_per_env_step"], 1),
                                "flop_per_env_step": round(counts[4096]["flop_per_env_step"]),
                                "dram_bytes_per_launch_4096": round(counts[4096]["dram_bytes_per_launch"]),
                                "source": f"profiles/{tag}_step_kernel_4096_ncu_raw.txt (ncu --metrics pass, mean over the captured launches)"}}
        if 65536 in counts:
            kc["walk_ik_plane_65536"] = {"warp_inst_per_env_step": round(counts[65536]["warp_inst_per_env_step"], 1),
                                         "flop_per_env_step": round(counts[65536]["flop_per_env_step"]),
                                         "dram_bytes_per_launch": round(counts[65536]["dram_bytes_per_launch"]),
                                         "source": f"profiles/{tag}_step_kernel_65536_ncu_raw.txt"}
        json.dump(kc, open(os.path.join(ROOT, "profiles", "kernel_counts.json"), "w"), indent=1)
        print("wrote profiles/kernel_counts.json", kc)


if __name__ == "__main__":
    main()
