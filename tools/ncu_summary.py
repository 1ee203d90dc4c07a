"""Print the metrics we quote from an .ncu-rep (raw page), one block per captured launch."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'sm__cycles_elapsed.avg', 'lts__t_sector_hit_rate.pct']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')][:60])
    for w in want:
        if w in hdr:
            print(f"  {w}: {r[hdr.index(w)]} {units[hdr.index(w)]}")
    top = sorted(((float(r[hdr.index(s)] or 0), s) for s in stalls), reverse=True)[:8]
    for v, s in top:
        print(f"  stall {s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")
