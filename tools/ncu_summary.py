"""summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'launch__grid_size', 'launch__block_size']
for r in rows[2:]:
    print("kernel:", r[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr:
            print(f"   {w:75s} {r[hdr.index(w)]} {units[hdr.index(w)]}")
    st = [(float(r[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
          for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
    print("   stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
