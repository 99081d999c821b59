"""Key metrics of every kernel in an ncu report, as text for profiles/:  python tools/ncu_summary.py <report.ncu-rep>"""
import csv, io, subprocess, sys

KEYS = [("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.avg.per_second", "SM clock"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs/SM)"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/smem throughput %"),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (tcgen05) active %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM ld/st pipe %"),
        ("sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "TMA pipe %")]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
print(f"# {rep}: ncu --set full --clock-control none (cold caches, one launch each; compare with the CUDA-event timings in bench.py)")
for r in rows[2:]:
    print(f"\n{r[col['Kernel Name']][:150]}")
    for k, name in KEYS:
        if k in col:
            print(f"    {name:34s} {r[col[k]]:>16s} {units[col[k]]}")
