#!/usr/bin/env python3
"""Generate the committed golden fixtures from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, read-only).  Nothing
here is imported by the product; tests read the files this script writes.

What it does
  1. copies the reference's example inputs / helper files (test/*.txt) and its
     seven shipped golden scans (test/output/*.txt) into tests/golden/data and
     tests/golden/ref_out  -- data fixtures, not source.
  2. builds two small synthetic inputs (mixed sample sizes; duplicate genetic
     positions) with a fixed seed.
  3. runs `python /root/reference/BalLeRMix+_v1.py ...` as a subprocess for every
     case in CASES (flags no shipped golden covers: B0, --findBal, --fixX,
     --fixAlpha, --listA, -w radius mode, float steps, --noCenter,
     --fixWinSize site-centred, --usePhysPos/--rec, mixed n, --getSpect /
     --getConfig) and stores the outputs under tests/golden/gen/.
  4. writes tests/golden/manifest.json (name -> argv, relative to tests/golden).

Usage:  python tests/golden/make_golden.py [-j 8] [--only NAME ...]
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
REF_SCRIPT = os.path.join(REF, "BalLeRMix+_v1.py")
DATA = os.path.join(HERE, "data")
REF_OUT = os.path.join(HERE, "ref_out")
GEN = os.path.join(HERE, "gen")

D = "data/"
EX1 = D + "Example1_fullSweep_200kya_DAF.txt"
EX1M = D + "Example1_fullSweep_200kya_MAF.txt"
EX2 = D + "Example2_balancing_10MYA_DAF.txt"
EX2NS = D + "Example2_balancing_10MYA_DAF_nosub.txt"
EX2M = D + "Example2_balancing_10MYA_MAF.txt"
EX2MNS = D + "Example2_balancing_10MYA_MAF_nosub.txt"
SP_B2 = D + "HC_CEU_Neut_DAF_spect_for_B2.txt"
SP_B2M = D + "HC_CEU_Neut_MAF_spect_for_B2maf.txt"
SP_B0 = D + "HC_CEU_Neut_DAF-nosub_spect_for_B0.txt"
SP_B0M = D + "HC_CEU_Neut_MAF-noSub_spect_for_B0maf.txt"
CF_B1 = D + "HC_CEU_Neut_config_for_B1.txt"
MIXN = D + "synth_mixed_n.txt"
MIXN_SP = "gen/synth_mixed_n_spect.txt"
DUPS = D + "synth_dup_genpos.txt"
N200 = D + "synth_n200.txt"
N200_SP = "gen/synth_n200_spect.txt"
LIST_A_100 = ",".join(str(float(a)) for a in range(1000, 11000, 100))     # == --rangeA 1000,10900,100

# name -> (argv after the script name, output file relative to tests/golden)
# scan cases write -o gen/<name>.txt ; helper cases write --spect gen/<name>.txt
CASES = {
    # --- helper-file generation (a7) ---
    "spect_ex1_B2": (["-i", EX1, "--getSpect", "--spect", "gen/spect_ex1_B2.txt"], "gen/spect_ex1_B2.txt"),
    "spect_ex1_B2maf": (["-i", EX1, "--getSpect", "--MAF", "--spect", "gen/spect_ex1_B2maf.txt"], "gen/spect_ex1_B2maf.txt"),
    "spect_ex2_B0": (["-i", EX2, "--getSpect", "--noSub", "--spect", "gen/spect_ex2_B0.txt"], "gen/spect_ex2_B0.txt"),
    "spect_ex2_B0maf": (["-i", EX2M, "--getSpect", "--noSub", "--MAF", "--spect", "gen/spect_ex2_B0maf.txt"], "gen/spect_ex2_B0maf.txt"),
    "spect_ex1M_B2maf": (["-i", EX1M, "--getSpect", "--MAF", "--spect", "gen/spect_ex1M_B2maf.txt"], "gen/spect_ex1M_B2maf.txt"),
    "config_ex1_B1": (["-i", EX1, "--getConfig", "--spect", "gen/config_ex1_B1.txt"], "gen/config_ex1_B1.txt"),
    "config_ex2_B1": (["-i", EX2, "--getConfig", "--spect", "gen/config_ex2_B1.txt"], "gen/config_ex2_B1.txt"),
    "synth_mixed_n_spect": (["-i", MIXN, "--getSpect", "--spect", MIXN_SP], MIXN_SP),
    "synth_n200_spect": (["-i", N200, "--getSpect", "--spect", N200_SP], N200_SP),
    # --- scans not pinned by a shipped golden ---
    "ex2_B0_s5": (["-i", EX2NS, "--spect", SP_B0, "--noSub", "-s", "5", "-o", "gen/ex2_B0_s5.txt"], "gen/ex2_B0_s5.txt"),
    "ex2_B0_dropsub_s20": (["-i", EX2, "--spect", SP_B0, "--noSub", "-s", "20", "-o", "gen/ex2_B0_dropsub_s20.txt"], "gen/ex2_B0_dropsub_s20.txt"),
    "ex2_B2maf_findBal_s60": (["-i", EX2M, "--spect", SP_B2M, "--MAF", "--findBal", "-s", "60", "-o", "gen/ex2_B2maf_findBal_s60.txt"], "gen/ex2_B2maf_findBal_s60.txt"),
    "ex1_B2_fixgrid": (["-i", EX1, "--spect", SP_B2, "--fixX", "0.3", "--fixAlpha", "20", "--listA", "100,1000,10000", "-o", "gen/ex1_B2_fixgrid.txt"], "gen/ex1_B2_fixgrid.txt"),
    "ex1_B2_w20_s10": (["-i", EX1, "--spect", SP_B2, "-w", "20", "-s", "10", "-o", "gen/ex1_B2_w20_s10.txt"], "gen/ex1_B2_w20_s10.txt"),
    "ex1_B2_w15_s7p5": (["-i", EX1, "--spect", SP_B2, "-w", "15", "-s", "7.5", "--listA", "100,500,2000,1e4,1e6", "-o", "gen/ex1_B2_w15_s7p5.txt"], "gen/ex1_B2_w15_s7p5.txt"),
    "ex2_B0maf_noCenter": (["-i", EX2MNS, "--spect", SP_B0M, "--noSub", "--MAF", "--usePhysPos", "--fixWinSize", "-w", "2000", "--noCenter", "-s", "1000", "-o", "gen/ex2_B0maf_noCenter.txt"], "gen/ex2_B0maf_noCenter.txt"),
    "ex2_B0maf_noCenter_gaps": (["-i", EX2MNS, "--spect", SP_B0M, "--noSub", "--MAF", "--usePhysPos", "--fixWinSize", "-w", "300", "--noCenter", "-s", "150", "--fixX", "0.5", "--listA", "100,1000,1e4", "-o", "gen/ex2_B0maf_noCenter_gaps.txt"], "gen/ex2_B0maf_noCenter_gaps.txt"),
    "ex1_B2_phys_s50": (["-i", EX1, "--spect", SP_B2, "--usePhysPos", "--rec", "1e-6", "-s", "50", "-o", "gen/ex1_B2_phys_s50.txt"], "gen/ex1_B2_phys_s50.txt"),
    "ex2_B2_fixwin_s50": (["-i", EX2, "--spect", SP_B2, "--usePhysPos", "--fixWinSize", "-w", "5000", "-s", "50", "-o", "gen/ex2_B2_fixwin_s50.txt"], "gen/ex2_B2_fixwin_s50.txt"),
    "ex2_B2_fixwin_nophys_s40": (["-i", EX2, "--spect", SP_B2, "--fixWinSize", "-w", "3000", "-s", "40", "--findBal", "--fixX", "0.4", "-o", "gen/ex2_B2_fixwin_nophys_s40.txt"], "gen/ex2_B2_fixwin_nophys_s40.txt"),
    "ex2_B1_listA_s10": (["-i", EX2, "--spect", CF_B1, "--noFreq", "--listA", "50,500,5e3", "--fixX", "0.5", "-s", "10", "-o", "gen/ex2_B1_listA_s10.txt"], "gen/ex2_B1_listA_s10.txt"),
    "ex2_B2_findBal_fixX_s100": (["-i", EX2, "--spect", SP_B2, "--findBal", "--fixX", "0.25", "-s", "100", "-o", "gen/ex2_B2_findBal_fixX_s100.txt"], "gen/ex2_B2_findBal_fixX_s100.txt"),
    "ex1M_B2maf_s80": (["-i", EX1M, "--spect", SP_B2M, "--MAF", "-s", "80", "-o", "gen/ex1M_B2maf_s80.txt"], "gen/ex1M_B2maf_s80.txt"),
    "ex1_B1_s80": (["-i", EX1, "--spect", CF_B1, "--noFreq", "-s", "80", "-o", "gen/ex1_B1_s80.txt"], "gen/ex1_B1_s80.txt"),
    "synth_mixed_n_B2_s20": (["-i", MIXN, "--spect", MIXN_SP, "-s", "20", "-o", "gen/synth_mixed_n_B2_s20.txt"], "gen/synth_mixed_n_B2_s20.txt"),
    # the bench workload's shape (n = 200, 100-point A grid, --usePhysPos --rec 1e-8) at a size the
    # reference finishes in minutes: 5 centres x 51 000 grid points
    "synth_n200_B2_s700": (["-i", N200, "--spect", N200_SP, "--usePhysPos", "--rec", "1e-8", "--listA", LIST_A_100,
                             "-s", "700", "-o", "gen/synth_n200_B2_s700.txt"], "gen/synth_n200_B2_s700.txt"),
    "synth_dup_genpos_B2": (["-i", DUPS, "--spect", SP_B2, "--fixX", "0.2", "--listA", "100,1000,1e6,1e8", "-o", "gen/synth_dup_genpos_B2.txt"], "gen/synth_dup_genpos_B2.txt"),
}
# cases whose inputs are produced by another case
DEPENDS = {"synth_mixed_n_B2_s20": "synth_mixed_n_spect", "synth_n200_B2_s700": "synth_n200_spect"}


def copy_reference_data():
    os.makedirs(DATA, exist_ok=True)
    os.makedirs(REF_OUT, exist_ok=True)
    os.makedirs(GEN, exist_ok=True)
    for f in sorted(os.listdir(os.path.join(REF, "test"))):
        if f.endswith(".txt"):
            shutil.copyfile(os.path.join(REF, "test", f), os.path.join(DATA, f))
    for f in sorted(os.listdir(os.path.join(REF, "test", "output"))):
        if f.endswith(".txt"):
            shutil.copyfile(os.path.join(REF, "test", "output", f), os.path.join(REF_OUT, f))


def make_synthetic_inputs():
    rng = np.random.default_rng(20261018)
    # (1) mixed sample sizes n in {48, 50}; DAF, substitutions included
    n_sites = 240
    pos = np.sort(rng.choice(np.arange(10, 60000), size=n_sites, replace=False))
    n = rng.choice([48, 50], size=n_sites, p=[0.35, 0.65])
    is_sub = rng.random(n_sites) < 0.6
    k = np.where(is_sub, n, np.minimum(n - 1, 1 + rng.geometric(0.12, size=n_sites)))
    with open(os.path.join(HERE, MIXN), "w") as fh:
        fh.write("physPos\tgenPos\tx\tn\n")
        for p, kk, nn in zip(pos, k, n):
            fh.write(f"{p}\t{float(p * 1e-6)!r}\t{kk}\t{nn}\n")
    # (2) duplicate genetic positions (several sites share genPos with the centre)
    #     counts restricted to classes the shipped B2 spectrum knows (n = 50)
    n_sites = 120
    pos = np.sort(rng.choice(np.arange(100, 20000), size=n_sites, replace=False))
    gen = np.round(pos * 1e-6, 4)  # coarse rounding -> ties in genPos
    is_sub = rng.random(n_sites) < 0.6
    k = np.where(is_sub, 50, rng.integers(1, 50, size=n_sites))
    with open(os.path.join(HERE, DUPS), "w") as fh:
        fh.write("physPos\tgenPos\tx\tn\n")
        for p, g, kk in zip(pos, gen, k):
            fh.write(f"{p}\t{float(g)!r}\t{kk}\t50\n")


def make_n200_input():
    """The bench generator (n = 200, P(sub) = 0.7, P(k) ~ 1/k, 70 nt spacing), 3000 sites."""
    rng = np.random.default_rng(777)
    n_sites, n = 3000, 200
    pos = np.sort(rng.choice(n_sites * 70 - 1, size=n_sites, replace=False)) + 1
    is_sub = rng.random(n_sites) < 0.7
    w = 1. / np.arange(1, n)
    k = np.where(is_sub, n, rng.choice(np.arange(1, n), size=n_sites, p=w / w.sum()))
    # a footprint of balancing selection around site 1400 (otherwise every row is the all-zero row):
    # within 6 kb most sites become polymorphisms at intermediate frequency
    near = np.abs(pos - pos[1400]) < 6000
    k = np.where(near & (rng.random(n_sites) < 0.8), rng.integers(70, 131, size=n_sites), k)
    with open(os.path.join(HERE, N200), "w") as fh:
        fh.write("physPos\tgenPos\tx\tn\n")
        for p, kk in zip(pos, k):
            fh.write(f"{p}\t{float(p * 1e-8)!r}\t{kk}\t{n}\n")


def run_case(name):
    argv, out = CASES[name]
    t0 = time.time()
    proc = subprocess.run([sys.executable, REF_SCRIPT] + argv, cwd=HERE,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    ok = proc.returncode == 0 and os.path.exists(os.path.join(HERE, out))
    return name, ok, time.time() - t0, proc.stdout[-400:] if not ok else ""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    ap.add_argument("--only", nargs="*")
    opt = ap.parse_args()
    copy_reference_data()
    make_synthetic_inputs()
    make_n200_input()
    names = opt.only or list(CASES)
    first = [n for n in names if n in DEPENDS.values()]
    rest = [n for n in names if n not in first]
    results = []
    for group in (first, rest):
        with ThreadPoolExecutor(max_workers=opt.j) as ex:
            for name, ok, dt, tail in ex.map(run_case, group):
                print(f"{'ok  ' if ok else 'FAIL'} {name:32s} {dt:7.1f}s {tail}", flush=True)
                results.append((name, ok))
    manifest = {n: {"argv": CASES[n][0], "output": CASES[n][1]} for n in CASES}
    with open(os.path.join(HERE, "manifest.json"), "w") as fh:
        json.dump({"reference": "bioXiaoheng/BallerMixPlus BalLeRMix+_v1.py (unmodified)",
                   "python": sys.version.split()[0], "numpy": np.__version__,
                   "scipy": __import__("scipy").__version__, "cases": manifest}, fh, indent=1)
    if not all(ok for _, ok in results):
        sys.exit(1)


if __name__ == "__main__":
    main()
