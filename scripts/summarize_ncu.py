"""Turn an .ncu-rep (read here, no GPU needed) into the short text summary committed under profiles/."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "local-load L1 hit %"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local load instr"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
]


def main(path, voxels=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    out = [f"# ncu --set full summary of {path.split('/')[-1]}", ""]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        out.append(f"## {name[:110]}")
        vals = {}
        for key, label in KEYS:
            if key in h:
                i = h.index(key)
                vals[key] = r[i]
                out.append(f"{label:45s} {r[i]} {units[i]}")
        try:
            dur = float(vals["gpu__time_duration.sum"])
            unit = units[h.index("gpu__time_duration.sum")]
            sec = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit, 1e-9)
            rd, wr = float(vals["dram__bytes_read.sum"]), float(vals["dram__bytes_write.sum"])
            ur = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            b = rd * ur.get(units[h.index("dram__bytes_read.sum")], 1) + wr * ur.get(units[h.index("dram__bytes_write.sum")], 1)
            out.append(f"{'dram bytes per launch (read+write)':45s} {b:.4e} B  -> {b / sec / 1e9:.1f} GB/s")
            if voxels:
                out.append(f"{'per voxel':45s} {b / voxels:.1f} B dram, {float(vals['smsp__inst_executed.sum']) / voxels:.0f} warp-instr")
        except (KeyError, ValueError):
            pass
        out.append("")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
