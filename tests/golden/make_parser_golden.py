#!/usr/bin/env python3
"""Fixtures and goldens for the upstream parser, made in the BUILD CONTAINER from the
unmodified reference (/root/reference/parsing_scripts/parse_ballermix_input.py) and committed:

    python tests/golden/make_parser_golden.py

  parser/shipped_first2000.vcf.gz, shipped_ids.txt, shipped_recmap.txt
        the reference's own Example3 fixtures, verbatim
  parser/shipped_vcf-only_rec1.25e-6.txt
        the reference's own expected output for its vcf-only example, verbatim (the script as shipped raises on
        that example -- ValueError('Invalid x: 0') -- so this file is the only statement of the intended output)
  parser/poly.vcf.gz
        the records of the shipped VCF that are polymorphic among the chosen samples (so that the reference does
        not raise), 130 sample columns
  parser/synthetic.axt
        a synthetic pairwise alignment of chr22 over the region of the VCF: three blocks, gaps in both
        sequences, N runs, lower-case bases, substitutions, and the VCF's REF or ALT as the outgroup base
  parser/ref_*.txt
        outputs of the reference for its four modes on these inputs
"""
import gzip
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, 'parser')
REF = '/root/reference/parsing_scripts'
VCF = os.path.join(REF, 'Example3_first2000var.chr22.phase3_shapeit2_mvncall_integrated_v5b.20130502.genotypes.vcf.gz')


def main():
    os.makedirs(OUT, exist_ok=True)
    shutil.copyfile(VCF, os.path.join(OUT, 'shipped_first2000.vcf.gz'))
    shutil.copyfile(os.path.join(REF, 'Example3_YRI_samples_1KG-v3.20130502.txt'), os.path.join(OUT, 'shipped_ids.txt'))
    shutil.copyfile(os.path.join(REF, 'Example3_YRI_0-1622e4_rec_map_hapmap_format_hg19_chr22.txt'),
                    os.path.join(OUT, 'shipped_recmap.txt'))
    shutil.copyfile(os.path.join(REF, 'test_output', 'Example3_vcf-only_rec1.25e-6_b0maf-ready.txt'),
                    os.path.join(OUT, 'shipped_vcf-only_rec1.25e-6.txt'))
    ids = open(os.path.join(OUT, 'shipped_ids.txt')).read().strip().split(',')

    # ---- poly.vcf.gz: 130 columns (the listed samples + 22 others), records polymorphic among the listed ones
    rng = np.random.default_rng(3)
    with gzip.open(VCF, 'rt') as fh:
        lines = fh.read().splitlines()
    meta = [ln for ln in lines if ln.startswith('##')][:5]
    head = next(ln for ln in lines if ln.startswith('#CHROM')).split('\t')
    listed = [head.index(i) for i in ids]
    others = [c for c in range(9, len(head)) if c not in set(listed)]
    keep = sorted(listed + rng.choice(others, size=22, replace=False).tolist())
    recs, kept_pos = [], []
    for ln in lines:
        if ln.startswith('#'):
            continue
        f = ln.split('\t')
        x = sum(f[c].count('1') for c in listed)
        n = 2 * len(listed)
        poly = f[3] in 'ACGT' and f[4] in 'ACGT' and len(f[3]) == 1 and len(f[4]) == 1 and 'PASS' in f[6]
        if poly and not (0 < x < n):
            continue
        recs.append('\t'.join(f[:9] + [f[c] for c in keep]))
        kept_pos.append((int(f[1]), f[3], f[4]))
    with gzip.open(os.path.join(OUT, 'poly.vcf.gz'), 'wt') as fh:
        fh.write('\n'.join(meta + ['\t'.join(head[:9] + [head[c] for c in keep])] + recs) + '\n')

    # ---- synthetic.axt over the region of the VCF
    lo, hi = kept_pos[0][0] - 40, kept_pos[-1][0] + 60
    snp = {p: (r, a) for p, r, a in kept_pos if len(r) == 1 and len(a) == 1}
    cuts = [lo, lo + (hi - lo) // 3, lo + 2 * (hi - lo) // 3 - 500, hi]      # three blocks, a hole before the third
    starts = [cuts[0], cuts[1], cuts[2] + 500]
    ends = [cuts[1] - 1, cuts[2] - 1, cuts[3]]
    blocks = []
    for b, (s, e) in enumerate(zip(starts, ends)):
        prim, alig = [], []
        p = s
        while p <= e:
            u = rng.random()
            base = 'ACGT'[rng.integers(4)]
            if p in snp:
                r, a = snp[p]
                base = r
                v = rng.random()
                out = r if v < 0.45 else a if v < 0.9 else 'ACGT'[rng.integers(4)] if v < 0.97 else '-'
            elif u < 0.004:
                out = 'ACGT'[(('ACGT'.index(base)) + 1 + rng.integers(3)) % 4]        # substitution
            elif u < 0.006:
                out = '-'                                                             # gap in the outgroup
            elif u < 0.007:
                out = 'N'
            else:
                out = base
            if rng.random() < 0.002 and p not in snp:                                 # insertion in the outgroup
                prim.append('-'); alig.append('ACGT'[rng.integers(4)])
            if rng.random() < 0.0005 and p not in snp:
                base = 'N'                                                            # N in the primary sequence
            lower = rng.random() < 0.3
            prim.append(base.lower() if lower else base)
            alig.append(out.lower() if lower else out)
            p += 1
        blocks.append(f'{b} chr22 {s} {e} chr22_out {s + 1000} {e + 1000} + {5000 + b}\n{"".join(prim)}\n{"".join(alig)}\n')
    with open(os.path.join(OUT, 'synthetic.axt'), 'w') as fh:
        fh.write('##matrix=axtChain 16 91,-114,-31,-123\n##gapPenalties=axtChain O=400 E=30\n' + '\n'.join(blocks) + '\n')

    # ---- the reference's outputs
    runs = {
        'ref_vcf-only.txt': ['--vcf', 'poly.vcf.gz', '--ID_list', 'shipped_ids.txt'],
        'ref_vcf-only_rate.txt': ['--vcf', 'poly.vcf.gz', '--ID_list', 'shipped_ids.txt', '--rec_rate', '1.25e-6'],
        'ref_vcf-recmap.txt': ['--vcf', 'poly.vcf.gz', '--ID_list', 'shipped_ids.txt', '--rec_map', 'shipped_recmap.txt'],
        'ref_vcf-recmap_all-samples.txt': ['--vcf', 'poly.vcf.gz', '--rec_map', 'shipped_recmap.txt', '--rec_rate', '2e-6'],
        'ref_vcf-axt.txt': ['--vcf', 'poly.vcf.gz', '--ID_list', 'shipped_ids.txt', '--axt', 'synthetic.axt'],
        'ref_vcf-axt_rate_hap.txt': ['--vcf', 'poly.vcf.gz', '--axt', 'synthetic.axt', '--rec_rate', '1.25e-6', '--hap'],
        'ref_vcf-axt-recmap.txt': ['--vcf', 'poly.vcf.gz', '--ID_list', 'shipped_ids.txt', '--axt', 'synthetic.axt',
                                   '--rec_map', 'shipped_recmap.txt'],
        'ref_vcf-axt-recmap_rate.txt': ['--vcf', 'shipped_first2000.vcf.gz', '--ID_list', 'shipped_ids.txt', '--axt',
                                        'synthetic.axt', '--rec_map', 'shipped_recmap.txt', '--rec_rate', '1.25e-6'],
    }
    for name, argv in runs.items():
        cmd = [sys.executable, os.path.join(REF, 'parse_ballermix_input.py'), '-c', '22', '-o', name] + argv
        res = subprocess.run(cmd, cwd=OUT, capture_output=True, text=True)
        rows = len(open(os.path.join(OUT, name)).read().splitlines()) if os.path.exists(os.path.join(OUT, name)) else -1
        print(name, 'rc', res.returncode, 'rows', rows, res.stderr.strip().splitlines()[-1:] if res.returncode else '')
    with open(os.path.join(OUT, 'manifest.txt'), 'w') as fh:
        for name, argv in runs.items():
            fh.write(name + '\t' + ' '.join(argv) + '\n')


if __name__ == '__main__':
    main()
