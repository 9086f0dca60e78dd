"""Build profiles/ncu_traffic.json: per C-ABI call of one training step, the device time and the DRAM traffic
(dram__bytes_read.sum + dram__bytes_write.sum) that ncu measured for its main kernel.

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-step
    python profiles/make_traffic_json.py gpurun_out/launches.csv gpurun_out/profile_step_calls.json

bench.py --profile-step writes the ordered list of C-ABI calls; the i-th launch of a main kernel in ncu's launch
list belongs to the i-th call of the matching entry point (finalize kernels are skipped)."""
import csv
import json
import os
import re
import sys
from collections import defaultdict

MAIN = {'cost_volume_fwd': 'uof_cost_volume_fwd', 'cost_volume_bwd': 'uof_cost_volume_bwd', 'warp_fwd': 'uof_warp_fwd',
        'warp_bwd': 'uof_warp_bwd', 'photo_loss_fwd_kernel': 'uof_photo_loss_fwd', 'photo_loss_bwd': 'uof_photo_loss_bwd',
        'photo_warp_fwd': 'uof_photo_warp_loss_fwd', 'photo_warp_bwd': 'uof_photo_warp_loss_bwd',
        'upsample_fwd': 'uof_upsample_bilinear_fwd', 'upsample_bwd': 'uof_upsample_bilinear_bwd',
        'smooth_fwd': 'uof_smooth_loss_fwd', 'smooth_bwd': 'uof_smooth_loss_bwd', 'consis_fwd': 'uof_consis_loss_fwd',
        'consis_bwd': 'uof_consis_loss_bwd', 'pyramid': 'uof_img_pyramid', 'ssim_fwd': 'uof_ssim_fwd', 'ssim_bwd': 'uof_ssim_bwd',
        'splat': 'uof_splat_fwd', 'bias_lrelu_fwd': 'uof_bias_lrelu_fwd', 'bias_lrelu_bwd': 'uof_bias_lrelu_bwd'}      # bench.py keys uof_bias_lrelu_bwd2 calls as uof_bias_lrelu_bwd[...]


def main(launch_csv, calls_json, out_path):
    calls = json.load(open(calls_json))
    by_entry = defaultdict(list)
    for k in calls:
        by_entry[k.split('[')[0]].append(k)
    with open(launch_csv) as f:
        lines = [l for l in f if not l.startswith('==')]
    recs, order = defaultdict(dict), []
    for row in csv.DictReader(lines):
        i = row['ID']
        if i not in recs:
            order.append(i)
            recs[i]['name'] = row['Kernel Name']
        v = float(row['Metric Value'].replace(',', ''))
        u, m = row['Metric Unit'], row['Metric Name']
        if m == 'gpu__time_duration.sum':
            v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
        if m.startswith('dram__bytes'):
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        recs[i][m] = v
    seen = defaultdict(int)
    table = {}
    for i in order:
        r = recs[i]
        if 'finalize' in r['name']:
            continue
        fn = re.sub(r'[<(].*', '', r['name'].replace('<unnamed>', 'anon')).split('::')[-1].strip()
        entry = next((e for k, e in MAIN.items() if fn.startswith(k)), None)
        if entry is None or seen[entry] >= len(by_entry[entry]):
            continue
        key = by_entry[entry][seen[entry]]
        seen[entry] += 1
        table[key] = {'kernel': fn, 'us': round(r.get('gpu__time_duration.sum', 0.0), 2),
                      'dram_bytes': int(r.get('dram__bytes_read.sum', 0) + r.get('dram__bytes_write.sum', 0))}
    json.dump(table, open(out_path, 'w'), indent=1, sort_keys=True)
    print('wrote %s (%d calls)' % (out_path, len(table)))


if __name__ == '__main__':
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ncu_traffic.json')
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else out)
