"""Text summary of an ncu --set full capture (read here, no GPU): the metrics B200_PROFILING.md names + stall reasons.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/r02_x_ncu_summary.txt
"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__t_sector_hit_rate\.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|launch__(block_size|grid_size|registers_per_thread|"
    r"shared_mem_per_block_dynamic|occupancy_limit_\w+)|sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
    r"sm__pipe_fma_cycles_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_(xu|lsu|alu|tmem|tma)\.avg\.pct_of_peak_sustained_active|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__inst_executed\.sum|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|sm__cycles_elapsed\.avg)$")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("kernel:", name[:150])
    for h, v, u in zip(hdr, r, units):
        if want.match(h):
            print(f"  {h:100s} {v:>18s} {u}")
