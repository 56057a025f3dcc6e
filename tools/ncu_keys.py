"""Print the key metrics of every kernel in an ncu report: python tools/ncu_keys.py rep.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg']
for vals in rows[2:]:
    print('---')
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w} = {vals[i]} {units[i]}")
    st = []
    for i, h in enumerate(hdr):
        if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct') and 'not_issued' not in h:
            try:
                v = float(vals[i])
                if v > 1.0: st.append((v, h.split('issue_stalled_')[1].replace('_per_warp_active.pct', '')))
            except Exception:
                pass
    print("  stalls(% of warp-active):", sorted(st, reverse=True)[:8])
