"""Text summary of one kernel of an .ncu-rep (the metrics DESIGN.md cites), for profiles/.
usage: python tools/ncu_summary.py report.ncu-rep [kernel-substring] > profiles/xxx.txt"""
import csv, subprocess, sys

KEYS = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__time_duration.sum', 'launch__block_size', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'sm__cycles_elapsed.avg', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__icc_request_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ''
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if want and want not in d.get('Kernel Name', ''):
            continue
        print('# kernel: %s  grid %s block %s' % (d.get('Kernel Name'), d.get('Grid Size'), d.get('Block Size')))
        u = dict(zip(hdr, units))
        for k in hdr:
            if k in KEYS or ('issue_stalled' in k and k.endswith('per_issue_active.ratio')):
                print('%-96s %-16s %s' % (k, u.get(k, ''), d[k]))
        break


if __name__ == '__main__':
    main()
