#!/usr/bin/env python3
"""Print the headline fields of a bench.py JSON line (last line of the given log)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value %.4g  ms/step %.1f' % (d['value'], d['ms_per_step']))
for k in ('e2e', 'strong', 'parity'):
    if k in d:
        v = d[k]
        print(k, {kk: v[kk] for kk in v if kk in ('value', 'ms_per_step', 'centres', 'rows_T_gt_0', 'max_rel_dT',
                                                  'argmax_mismatch', 'nsites_mismatch', 'ok', 'rows_equal_resident_run')})
r = d.get('roofline', {})
print('roofline', {k: r.get(k) for k in ('achieved', 'frac', 'kernel_ms_per_launch', 'kernel_share_of_step', 'sites_far_frac')})
print('work', r.get('work'))
if 'direct' in d:
    print('direct', d['direct']['value'], d['direct']['roofline']['frac'], d['direct']['roofline']['kernel_ms_per_launch'])
if 'cpu_baseline' in d:
    print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
