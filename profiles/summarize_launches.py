"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum [+ inst/dram metrics]) per own kernel launch.
Usage: python profiles/summarize_launches.py gpurun_out/rN_launches.csv"""
import csv
import re
import sys
from collections import defaultdict


OWN = re.compile(r'cost_volume_|warp_(fwd|bwd)_n|photo_loss_|photo_warp_|upsample_(fwd|bwd)(_int)?_kernel|smooth_(fwd|bwd|finalize)|consis_(fwd|bwd|finalize)|pyramid_|ssim_(fwd|bwd)_kernel|'
                 r'splat|fb_mask|clamp01|diff_weight_|masked_mean_|bias_lrelu_')


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    recs = defaultdict(dict)
    order = []
    for row in csv.DictReader(lines):
        key = row['ID']
        if key not in recs:
            order.append(key)
            recs[key].update(name=row['Kernel Name'], grid=row['Grid Size'], block=row['Block Size'])
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        m = row['Metric Name']
        if m == 'gpu__time_duration.sum':
            v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
        if m.startswith('dram__bytes'):
            v = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0) * v
        recs[key][m] = v
    tot = own = 0.0
    print('%-34s %-16s %9s %10s %9s %9s' % ('kernel', 'grid', 'us', 'warp-inst', 'rd MB', 'wr MB'))
    for k in order:
        r = recs[k]
        t = r.get('gpu__time_duration.sum', 0.0)
        tot += t
        if not OWN.search(r['name']):
            continue
        own += t
        nm = re.sub(r'\(.*', '', r['name']).replace('void ', '').replace('uof::', '').replace('cv::', '').replace('<unnamed>::', '')
        print('%-34s %-16s %9.1f %10.0f %9.2f %9.2f' % (nm[:34], r['grid'], t, r.get('smsp__inst_executed.sum', 0),
                                                          r.get('dram__bytes_read.sum', 0), r.get('dram__bytes_write.sum', 0)))
    print('own kernels %.1f us of %.1f us total (%.2f %%), %d launches in the step' % (own, tot, 100 * own / tot, len(order)))


if __name__ == '__main__':
    main(sys.argv[1])
