"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of numbers DESIGN.md / profiles/ quote.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {r[hdr.index('Kernel Name')]}  (ID {r[0]})")
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f"   {label:34s} {r[i]} {units[i]}")
        if "--stalls" in sys.argv:
            st = []
            for i, h in enumerate(hdr):
                if h.startswith(STALL_PREFIX) and h.endswith("_per_issue_active.ratio"):
                    try:
                        st.append((float(r[i]), h[len(STALL_PREFIX):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            for v, n in sorted(st, reverse=True)[:7]:
                print(f"   stalled warps per issue: {n:22s} {v:.2f}")
        print()


if __name__ == "__main__":
    main()
