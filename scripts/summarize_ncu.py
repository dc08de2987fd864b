"""Summarise an ncu report (.ncu-rep) into profiles/<name>.json + .txt (run on the build box).

    python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_scl8 --frames 75776 [--note "..."]
"""
import argparse
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_shared_mem_blocks",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc_per_sm",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_utilisation_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "avg_active_threads",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_pipe_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflict_wavefronts",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed": "dram_read_pct_of_peak",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_pct",
}
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "not_selected", "no_instruction", "math_pipe_throttle",
          "branch_resolving", "mio_throttle", "lg_throttle", "dispatch_stall", "barrier"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--frames", type=int, required=True)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"report": a.report, "frames_in_launch": a.frames, "note": a.note}
    for h, u, v in zip(hdr, units, vals):
        if h == "Kernel Name":
            d["kernel"] = v
        if h in KEYS:
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            d[KEYS[h]] = x * UNIT.get(u, 1.0) if KEYS[h] in ("duration", "dram_read", "dram_write", "l2_bytes") else x
        for st in STALLS:
            if h == f"smsp__average_warps_issue_stalled_{st}_per_issue_active.ratio":
                d.setdefault("stall_cycles_per_issue", {})[st] = float(v)
    F = a.frames
    d["per_frame"] = {
        "warp_instructions": d.get("warp_instructions", 0) / F,
        "dram_bytes": (d.get("dram_read", 0) + d.get("dram_write", 0)) / F,
        "smem_wavefronts": d.get("smem_wavefronts", 0) / F,
        "frames_per_second_under_ncu": F / d["duration"] if d.get("duration") else None,
    }
    with open(a.out + ".json", "w") as fh:
        json.dump(d, fh, indent=1)
    with open(a.out + ".txt", "w") as fh:
        fh.write(f"# {a.out}: {d.get('kernel')}\n# {a.note}\n")
        for k, v in d.items():
            if k not in ("kernel", "note"):
                fh.write(f"{k}: {v}\n")
    print(json.dumps(d["per_frame"]), d.get("ipc_per_sm"), d.get("issue_slot_utilisation_pct"))


if __name__ == "__main__":
    main()
