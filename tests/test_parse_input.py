"""Upstream parser: ballermixplus_b200.parse_input against outputs of the unmodified
reference parser (tests/golden/make_parser_golden.py, run in the build container) and against the reference's
own shipped expected output.  Byte-for-byte."""
import filecmp
import os

import pytest

import util
from ballermixplus_b200 import parse_input

PARSER = os.path.join(util.GOLD, 'parser')


def _manifest():
    with open(os.path.join(PARSER, 'manifest.txt')) as fh:
        return [ln.rstrip('\n').split('\t') for ln in fh if ln.strip()]


def _run(argv, out):
    full = []
    it = iter(argv)
    for a in it:
        full.append(a)
        if a in ('--vcf', '--ID_list', '--axt', '--rec_map'):
            full.append(os.path.join(PARSER, next(it)))
    with util.quiet():
        parse_input.main(['-c', '22', '-o', out] + full)


@pytest.mark.parametrize('name,argv', _manifest())
def test_matches_reference_output(name, argv, tmp_path):
    out = str(tmp_path / name)
    _run(argv.split(' '), out)
    assert filecmp.cmp(out, os.path.join(PARSER, name), shallow=False), name


def test_shipped_example_output(tmp_path):
    """The reference's README example (vcf only, --rec_rate 1.25e-6) on its own VCF and sample list: the
    reference as shipped raises on the first monomorphic site; its shipped expected output is the golden."""
    out = str(tmp_path / 'out.txt')
    _run(['--vcf', 'shipped_first2000.vcf.gz', '--ID_list', 'shipped_ids.txt', '--rec_rate', '1.25e-6'], out)
    assert filecmp.cmp(out, os.path.join(PARSER, 'shipped_vcf-only_rec1.25e-6.txt'), shallow=False)


def test_output_feeds_the_scan_reader(tmp_path):
    """The parser's table is what InputData reads (same four columns, header skipped)."""
    from ballermixplus_b200 import InputData
    out = str(tmp_path / 'out.txt')
    _run(['--vcf', 'poly.vcf.gz', '--ID_list', 'shipped_ids.txt', '--axt', 'synthetic.axt'], out)
    with util.quiet():
        data = InputData(out, False, False, False, 1)
    with open(out) as fh:
        rows = fh.read().splitlines()[1:]
    assert data.numSites == len(rows) and data.numSites > 900
    assert int(data.count.max()) == int(data.total.max()) == 216            # substitutions: x = n = 2 * 108


def test_errors_and_edge_cases(tmp_path):
    vcf = tmp_path / 'tiny.vcf'
    head = '##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\tS3\n'
    vcf.write_text(head
                   + '22\t100\t.\tA\tG\t1\tPASS\t.\tGT:DP\t0|1:3\t.|.:0\t1|1:2\n'      # missing genotype shortens n
                   + '22\t150\t.\tA\tGT\t1\tPASS\t.\tGT\t0|1\t0|0\t0|0\n'             # not a SNP
                   + '22\t170\t.\tC\tT\t1\tq10\t.\tGT\t0|1\t0|0\t0|0\n'               # filtered
                   + '21\t180\t.\tC\tT\t1\tPASS\t.\tGT\t0|1\t0|0\t0|0\n'              # other chromosome
                   + 'chr22\t200\t.\tC\tT\t1\tPASS\t.\tGT\t0|0\t0|0\t0|0\n'           # monomorphic: skipped
                   + 'chr22\t300\t.\tC\tT\t1\tPASS\t.\tGT\t0/1\t1\t0|0\n')            # haploid call mixes in
    out = str(tmp_path / 'o.txt')
    with util.quiet():
        parse_input.main(['--vcf', str(vcf), '-c', '22', '-o', out])
    assert open(out).read() == 'position\tgenPos\tx\tn\n100\t9.999999999999999e-05\t1\t4\n300\t0.0003\t2\t5\n'
    bad = tmp_path / 'tiny.txt'
    bad.write_text(head)
    with pytest.raises(SystemExit), util.quiet():
        parse_input.main(['--vcf', str(bad), '-c', '22', '-o', out])                 # not .vcf / .vcf.gz
    tri = tmp_path / 'tri.vcf'
    tri.write_text(head + '22\t100\t.\tA\tG\t1\tPASS\t.\tGT\t0|2\t0|0\t0|0\n')
    with pytest.raises(SystemExit), util.quiet():
        parse_input.main(['--vcf', str(tri), '-c', '22', '-o', out])                 # allele index 2
