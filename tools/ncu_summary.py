"""Key metrics from an .ncu-rep (`ncu -i rep --page raw --csv`) as one block per kernel launch."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'sm__cycles_active.avg', 'smsp__cycles_active.avg']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        print('==', r[ki][:110])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'   {w:72s} {r[i]:>16s} {units[i]}')
        for i, h in enumerate(hdr):            # tensor-pipe / TMEM / shared-memory operand-fetch counters, whatever this ncu calls them
            if h not in WANT and any(t in h for t in ('pipe_tensor', 'tc_wavefronts', 'tmem', 'data_pipe_lsu_wavefronts_mem_shared.sum',
                                                      'shared_op_ld.sum', 'achieved_occupancy')):
                print(f'   {h:72s} {r[i]:>16s} {units[i]}')


if __name__ == '__main__':
    main(sys.argv[1])
