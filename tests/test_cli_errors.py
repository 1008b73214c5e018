"""Error behaviour of the drop-in CLI (no GPU needed: every case stops before the scan).
The reference reports bad inputs with a message on stdout and sys.exit() -- status 0 -- and
so does the mirror (SURVEY.md §5 / §8b, v1:65-71, 192-194, 205-207, 218-221, 245-248, 286-287, 616-618, 756-758)."""
import contextlib
import io
import os

import pytest

import util
from ballermixplus_b200.cli import main

D = os.path.join(util.GOLD, 'data')
EX1 = os.path.join(D, 'Example1_fullSweep_200kya_DAF.txt')
SP_B2 = os.path.join(D, 'HC_CEU_Neut_DAF_spect_for_B2.txt')
SP_B0 = os.path.join(D, 'HC_CEU_Neut_DAF-nosub_spect_for_B0.txt')


def run(argv):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), pytest.raises(SystemExit) as exc:
        main(argv)
    return exc.value.code, buf.getvalue()


def test_no_arguments_prints_help_and_exits_zero():
    code, out = run([])
    assert code in (None, 0) and 'usage' in out.lower() and '--spect' in out


def test_zero_counts_in_daf_input(tmp_path):
    f = tmp_path / 'in.txt'
    f.write_text('physPos\tgenPos\tx\tn\n10\t1e-5\t0\t50\n20\t2e-5\t3\t50\n')
    code, out = run(['-i', str(f), '--spect', SP_B2, '-o', str(tmp_path / 'o.txt')])
    assert code in (None, 0) and 'zero derived alleles' in out
    assert not (tmp_path / 'o.txt').exists()


def test_data_class_missing_from_helper_file(tmp_path):
    f = tmp_path / 'in.txt'
    f.write_text('physPos\tgenPos\tx\tn\n10\t1e-5\t7\t48\n20\t2e-5\t3\t50\n')       # n = 48 is not in the spect
    code, out = run(['-i', str(f), '--spect', SP_B2, '-o', str(tmp_path / 'o.txt')])
    assert code in (None, 0) and 'not included in the helper file' in out


def test_spectrum_not_summing_to_one(tmp_path):
    sp = tmp_path / 'sp.txt'
    sp.write_text('1\t50\t0.4\n2\t50\t0.5\n')
    code, out = run(['-i', EX1, '--spect', str(sp), '-o', str(tmp_path / 'o.txt')])
    assert code in (None, 0) and 'do not add up to 1' in out and 'Sum = 0.9' in out


def test_config_must_sum_to_exactly_one(tmp_path):
    cf = tmp_path / 'cf.txt'
    cf.write_text('50\t0.7\t0.29999\n')
    code, out = run(['-i', EX1, '--spect', str(cf), '--noFreq', '-o', str(tmp_path / 'o.txt')])
    assert code in (None, 0) and 'do not add up to 1' in out


def test_nosub_spectrum_with_substitutions(tmp_path):
    code, out = run(['-i', EX1, '--spect', SP_B2, '--noSub', '-o', str(tmp_path / 'o.txt')])
    assert code in (None, 0) and 'Please do not account for substitutions' in out


def test_fixed_window_without_width(tmp_path):
    nosub = os.path.join(D, 'Example2_balancing_10MYA_DAF_nosub.txt')
    code, out = run(['-i', nosub, '--spect', SP_B0, '--noSub', '--fixWinSize', '-o', str(tmp_path / 'o.txt')])
    assert code in (None, 0) and 'Please set a window width' in out


def test_getspect_exits_after_writing(tmp_path):
    out_file = tmp_path / 'spect.txt'
    code, out = run(['-i', EX1, '--getSpect', '--spect', str(out_file)])
    assert code in (None, 0) and 'Done.' in out
    with open(os.path.join(util.GOLD, 'gen', 'spect_ex1_B2.txt')) as fh:
        assert out_file.read_text() == fh.read()
