"""Summarise an .ncu-rep (raw page) into a compact table: python tools/ncu_summary.py file.ncu-rep [out.md]"""
import csv
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'time'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__registers_per_thread', 'regs'), ('launch__shared_mem_per_block_dynamic', 'dsmem'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('smsp__inst_executed.sum', 'warp_inst'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
        ('sm__inst_executed_pipe_tensor.sum', 'tensor_inst'),
        ('l1tex__t_sector_hit_rate.pct', 'l1hit%'), ('lts__t_sector_hit_rate.pct', 'l2hit%'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_conflicts'),
        ('smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'st_long_sb'),
        ('smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'st_short_sb'),
        ('smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'st_barrier'),
        ('smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'st_mio'),
        ('smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'st_lg'),
        ('smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'st_math'),
        ('smsp__warp_issue_stalled_wait_per_warp_active.pct', 'st_wait'),
        ('smsp__warp_issue_stalled_not_selected_per_warp_active.pct', 'st_notsel')]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    seen = {}
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')].split('(')[0]
        seen[name] = seen.get(name, 0) + 1
        if seen[name] > 1:
            continue
        out.append('### `%s`' % name)
        cells = []
        for key, short in WANT:
            if key in hdr:
                v, u = r[hdr.index(key)], units[hdr.index(key)]
                try:
                    fv = float(v.replace(',', ''))
                    v = ('%.3g' % fv) if abs(fv) < 1e6 else ('%.4g' % fv)
                except ValueError:
                    pass
                cells.append('%s=%s%s' % (short, v, ('' if u in ('%', '') else ' ' + u)))
        out.append(', '.join(cells))
        out.append('')
    text = '\n'.join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'a').write(text + '\n')
    print(text)


if __name__ == '__main__':
    main()
