"""Prints the handful of ncu raw-page metrics we steer by.  usage: python tools/ncu_key.py raw.csv"""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__cycles_elapsed.max', 'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('----')
    for w in WANT:
        if w in hdr:
            print(f'{w} = {r[hdr.index(w)]} {units[hdr.index(w)]}')
    st = sorted(((float(r[hdr.index(h)].replace(',', '')), h) for h in stall), reverse=True)[:7]
    for v, h in st:
        print(f'   stall {h.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio","")}: {v:.2f}')
