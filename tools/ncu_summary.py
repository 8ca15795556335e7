"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few counters DESIGN.md / profiles/ cite."""
import csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.sum', 'smsp__thread_inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_fmalite.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_uniform.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_tmem_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tc.sum', 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg']


def main(path, extra=()):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==== %s  grid %s block %s' % (r[hdr.index('Kernel Name')][:90], r[hdr.index('Grid Size')] if 'Grid Size' in hdr else '', r[hdr.index('Block Size')] if 'Block Size' in hdr else ''))
        for k in list(KEYS) + [h for h in hdr if any(e in h for e in extra)]:
            if k in hdr:
                print('  %-78s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h]
        vals = sorted(((float(r[hdr.index(h)].replace(',', '') or 0), h) for h in st), reverse=True)[:7]
        print('  top stalls (warps per issue-active): ' + ', '.join('%s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for v, h in vals))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2:])
