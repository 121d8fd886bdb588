"""Summarise an `ncu --set full` report: one block per profiled launch with the metrics DESIGN.md / bench.py cite.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/raw.csv ; python scripts/ncu_summary.py /tmp/raw.csv
"""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active (% of peak, active cycles)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (% of peak)"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput (% of peak)"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput (% of peak)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of max warps)"),
    ("smsp__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe (% of peak)"),
    ("smsp__issue_active.avg.pct", "issue slots busy (%)"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    print(f"== {name}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
    for key, label in WANT:
        if key in idx and r[idx[key]] != "":
            print(f"   {label:52s} {r[idx[key]]:>16s} {units[idx[key]]}")
