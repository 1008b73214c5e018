#!/usr/bin/env python3
"""Turn the ncu outputs of tools/profile.sh into the committed summaries under profiles/.

    python tools/summarise_profile.py <tag>     # reads gpurun_out/<tag>_launches.csv, <tag>_prof.ncu-rep

Writes profiles/<tag>_launches.csv (the raw per-launch list), profiles/<tag>_launches_by_kernel.csv,
profiles/<tag>_scan_kernel_metrics.csv (selected raw-page metrics of the captured scan_kernel
launches), profiles/<tag>_summary.md, and profiles/traffic.json (DRAM bytes per launch, read by bench.py).
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_static', 'launch__waves_per_multiprocessor',
    'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
    'smsp__issue_active.avg.per_cycle_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
    'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed',
    'sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
]


def to_bytes(value, unit):
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return float(value) * scale.get(unit, 1)


def main():
    tag = sys.argv[1]
    src = os.path.join(ROOT, 'gpurun_out')
    dst = os.path.join(ROOT, 'profiles')
    os.makedirs(dst, exist_ok=True)
    out = [f'# ncu summary `{tag}`', '',
           'Produced by `tools/profile.sh` on a B200 (gpurun) and `tools/summarise_profile.py` here.',
           'Command profiled: `python bench.py --sites 10000000 --steps 1 --warmup 3 --profile ' + ' '.join(sys.argv[2:]) + '`'
           ' -- the benchmark\'s own configuration (10 M sites / 22 chromosomes, -s 1024); the captured launches are',
           'chromosome 1 (846 k sites, 827 centres x 100 A) and chromosome 2.',
           '(ncu launch times are cold-cache and serialised: compare shares, not absolutes).', '']

    # ---- launch list
    lpath = os.path.join(src, f'{tag}_launches.csv')
    if not os.path.exists(lpath):
        lpath = None
        out += ['(no launch list in this capture)', '']
    else:
        shutil.copyfile(lpath, os.path.join(dst, f'{tag}_launches.csv'))
    rows = [r for r in csv.reader(open(lpath)) if len(r) > 5] if lpath else [['Kernel Name', 'Metric Value']]
    hdr = rows[0]
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(',', ''))
        except ValueError:
            continue
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    own = sum(a[1] for k, a in agg.items() if 'dfma_peak' not in k) or 1.0
    with open(os.path.join(dst, f'{tag}_launches_by_kernel.csv') if lpath else os.devnull, 'w') as fh:
        fh.write('kernel,launches,total_ns,share_of_all,share_without_peak_probe\n')
        out += ['## Launch list (gpu__time_duration.sum)', '',
                '| kernel | launches | total ms | share | share w/o the FP64-peak probe |', '|---|---|---|---|---|']
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            share2 = '' if 'dfma_peak' in k else f'{100 * a[1] / own:.2f}%'
            fh.write(f'"{k}",{a[0]},{a[1]:.0f},{a[1] / tot:.5f},{share2}\n')
            out.append(f'| `{k[:90]}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.2f}% | {share2} |')
    out.append('')

    # ---- full capture of scan_kernel
    rep = os.path.join(src, f'{tag}_prof.ncu-rep')
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(dst, f'{tag}_scan_kernel_metrics.csv'), 'w') as fh:
        w = csv.writer(fh)
        w.writerow(['metric', 'unit'] + [f'launch{i}' for i in range(len(rows) - 2)])
        for k in ['Kernel Name'] + KEYS:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [r[i] for r in rows[2:]])
    r = rows[2]

    def get(k):
        return float(r[hdr.index(k)])

    dram = (to_bytes(r[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_read.sum')])
            + to_bytes(r[hdr.index('dram__bytes_write.sum')], units[hdr.index('dram__bytes_write.sum')]))
    dfma = get('smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed')
    dmul = get('smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed')
    dadd = get('smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed')
    peak = get('sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained')
    # algorithmic HBM bytes of the captured launch (chromosome 1): site arrays 20 B/site, the R table, the far-field
    # block + superblock moments of the chromosome, candidates written and re-read
    algo = None
    try:
        sys.path.insert(0, ROOT)
        import bench
        n1 = bench.genome_sizes(10_000_000)[0]
        items = len(range(0, n1, bench.BASE_STRIDE)) * 100
        algo = 20.0 * n1 + 8.0 * 512 * 200 + (n1 // 32 + n1 // 256) * 100 * 2 * 256.0 + 2 * 16.0 * items
    except Exception:
        pass
    out += ['## scan_kernel<16,4,far>, first captured launch (`ncu --set full`)', '',
            f'* duration {r[hdr.index("gpu__time_duration.sum")]} {units[hdr.index("gpu__time_duration.sum")]}, '
            f'grid {r[hdr.index("launch__grid_size")]} x {r[hdr.index("launch__block_size")]} threads, '
            f'{r[hdr.index("launch__registers_per_thread")]} registers/thread, '
            f'{r[hdr.index("launch__shared_mem_per_block_static")]} KB static smem/CTA',
            f'* FP64 pipe active: {get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"):.1f}% of '
            f'active cycles ({get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"):.1f}% of elapsed)',
            f'* executed FP64 thread-instructions per elapsed cycle: DFMA {dfma:.0f}, DMUL {dmul:.0f}, DADD {dadd:.0f} '
            f'(pipe peak {peak:.0f}/cycle) -> executed flops = 2*DFMA+DMUL+DADD = {2 * dfma + dmul + dadd:.0f}/cycle '
            f'= {100 * (2 * dfma + dmul + dadd) / (2 * peak):.1f}% of the DFMA flop peak',
            f'* warps: {get("smsp__warps_active.avg.per_cycle_active"):.2f} active, '
            f'{get("smsp__warps_eligible.avg.per_cycle_active"):.2f} eligible per scheduler cycle; '
            f'issue slots used {100 * get("smsp__issue_active.avg.per_cycle_active"):.0f}%',
            f'* DRAM traffic: {dram / 1e6:.2f} MB per launch '
            f'({get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):.3f}% of DRAM peak)'
            + (f'; algorithmic HBM bytes of this launch (site arrays + table + block moments + candidates) '
               f'{algo / 1e6:.0f} MB -> traffic / algorithmic = {dram / algo:.2f}' if algo else '')
            + ': the kernel is FP64-bound',
            '* top stall reasons (warps per issue): '
            + ', '.join(f'{k.split("stalled_")[1].split("_per_")[0]} {get(k):.2f}' for k in KEYS if 'issue_stalled' in k),
            '', f'All selected metrics for every captured launch: `{tag}_scan_kernel_metrics.csv`.', '']
    with open(os.path.join(dst, f'{tag}_summary.md'), 'w') as fh:
        fh.write('\n'.join(out))
    with open(os.path.join(dst, 'traffic.json'), 'w') as fh:
        json.dump({'dram_bytes_per_launch': dram, 'algorithmic_bytes_of_that_launch': algo,
                   'source': f'{tag}_prof.ncu-rep, scan_kernel<16,4,far>, launch of chromosome 1 (846 k sites), '
                   'dram__bytes_read.sum + dram__bytes_write.sum, bench.py --sites 10000000 --profile'}, fh)
    print('\n'.join(out))


if __name__ == '__main__':
    main()
