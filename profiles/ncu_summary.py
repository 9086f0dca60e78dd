"""Condense an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md quotes.
Usage: python profiles/ncu_summary.py report.ncu-rep [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'rd'), ('dram__bytes_write.sum', 'wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__registers_per_thread', 'regs'), ('smsp__inst_executed.sum', 'inst'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem_wf'), ('lts__t_sector_hit_rate.pct', 'l2hit%'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'st_long_sb'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'st_short_sb'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'st_barrier'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'st_math'),
        ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'st_mio'),
        ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'st_lg'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'st_wait'),
        ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'st_notsel'),
        ('smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio', 'st_sleep'),
        ('smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'st_dispatch')]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    for d in data:
        name = d[ix['Kernel Name']]
        if pat and not pat.search(name):
            continue
        short = re.sub(r'\(.*', '', name).replace('void ', '').replace('uof::<unnamed>::', '')
        out = ['%s grid=%s' % (short[:40], d[ix['Grid Size']].replace(' ', ''))]
        for m, label in WANT:
            if m in ix:
                v = d[ix[m]]
                u = units[ix[m]]
                try:
                    f = float(v.replace(',', ''))
                    v = '%.4g' % f
                except ValueError:
                    pass
                out.append('%s=%s%s' % (label, v, u if label in ('rd', 'wr') else ''))
        print('  '.join(out))


if __name__ == '__main__':
    main()
