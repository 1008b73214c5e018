"""Upstream parser -- SURVEY.md §8 row f4: VCF (+ AXT alignment, + HapMap-format recombination map) -> the
four-column input table of the scan (``position  genPos  x  n``).

Mirrors the command line of the reference's ``parsing_scripts/parse_ballermix_input.py`` ("ref", lines cited as
ref:N): same flags (``--vcf -c -o --ID_list --axt --rec_rate --rec_map --hap``), same four modes, same output
bytes.  The output is produced by the same arithmetic in the same order (positions and genetic positions are
formatted by Python, so every float prints as the reference prints it); the structure is different: one
allele counter, one genetic-map cursor and one alignment table shared by the four modes instead of four
copies of the loop.

Behaviour kept, quirks included (each one changes output bytes):
  * only bi-allelic SNPs (REF and ALT in ACGT) whose FILTER contains PASS are counted (ref:91,183,377,612);
  * genotypes: every integer in the GT field is an allele, '.' is missing and shortens n (ref:97-114);
  * vcf only: genPos = float(POS) * rec_rate, the position is echoed as the VCF wrote it (ref:117);
  * recombination map: the first line of the map is always skipped as a header (ref:135-137: the test
    compares a list with a string); a first segment with cumulative cM 0 has rate 0 (ref:141-142); positions
    past the last map row use ``--rec_rate`` up to 8e8 (ref:197-201); the two interpolation formulas and the
    state they update are the reference's (ref:187-218, 566-594, 659-684), including the use of the VCF
    position in the first formula for substitutions (ref:568);
  * alignment: positions advance only on A/C/G/T of the primary sequence, a pair is recorded when both
    bases are A/C/G/T (ref:305-317); substitutions (ref != outgroup) between two VCF records are written
    as x = n = ploidy * samples (ref:395-399, 563-595); polymorphisms are polarised with the four rules of
    ref:437-448; x == 0 (after polarisation) rows are dropped (ref:470-478);
  * with --axt and no map the FIRST VCF record is skipped (ref:361-362 read two lines);
  * when the alignment runs out, the rest of the VCF is dropped (ref:407-409, 600-601).

Where the reference cannot serve as the oracle:
  * without --axt the reference raises ``ValueError('Invalid x')`` on the first monomorphic site of the
    chosen samples (ref:119,247) -- its own shipped example stops there.  The outputs it ships
    (``test_output/Example3_vcf-only_rec1.25e-6_b0maf-ready.txt``) simply lack those rows, so such rows are
    skipped here; that shipped file is the golden for this mode.
  * an AXT file holding other chromosomes makes the reference index past a split sequence line
    (ref:288-297); such blocks are skipped here.  A VCF record of another chromosome under --axt loops
    forever in the reference (ref:384-389); here it is an error.
"""
import gzip
import re
import sys
import time

BASES = frozenset('ATCG')
_INTS = re.compile(r'[0-9]+')


def _open_text(filename, suffix=''):
    """ref:44-53"""
    if filename.lower().endswith('.gz'):
        return gzip.open(filename, 'rt')
    if filename.lower().endswith(suffix):
        return open(filename, 'r')
    print(f'Unrecognized {suffix.upper()} file name. Please make sure it\'s in {suffix} or {suffix}.gz format.')
    sys.exit(1)


def sample_columns(header_line, pop_list):
    """Columns of the VCF to count (ref:13-34).  With an ID list the reference keeps a *set* of column indices;
    counts do not depend on the order."""
    header = header_line.strip().split('\t')
    if pop_list is not None:
        with open(pop_list, 'r') as fh:
            ids = fh.read().strip().split(',')
        cols = set(map(header.index, ids))
        assert min(cols) >= 9
    else:
        cols = range(9, len(header))
    print(f'Data from {len(cols)} samples will be counted.')
    return cols


def count_alleles(fields, cols):
    """(x, n): ALT alleles and called alleles over the chosen samples (ref:93-114)."""
    gt = fields[8].split(':').index('GT')
    x = n = 0
    for i in cols:
        call = fields[i].split(':')[gt]
        alleles = tuple(map(int, _INTS.findall(call)))
        if len(alleles) == 0:
            assert '.' in call
            continue
        if max(alleles) > 1:
            print('Warning: This script only applies to diploid and haploid data.')
            print(call, alleles)
            sys.exit(1)
        x += sum(alleles)
        n += len(alleles)
    return x, n


def _is_counted_snp(fields):
    return fields[3] in BASES and fields[4] in BASES and 'PASS' in fields[6]


class MapCursor:
    """Genetic position (cM) of increasing physical positions from a HapMap-format map: chromosome, position,
    rate (cM/Mb), cumulative cM.  State and update rules are the reference's (ref:131-146, 187-218)."""

    def __init__(self, path, chrom, rec_rate):
        self.fh = _open_text(path)
        self.rec_rate = rec_rate
        row = self.fh.readline().strip().split('\t')
        if _INTS.findall(row[1]) != row[1]:            # always true (list vs str): the first line is dropped
            print('Skipping header.')
            row = self.fh.readline().strip().split('\t')
        assert row[0] in {chrom, 'chr' + chrom}
        self.r_pos, self.rate, self.cum = map(float, row[1:])
        if self.cum == 0:
            self.rate = 0
        self.last_r_pos, self.last_rate, self.last_cum = 0, self.rate, 0
        self.last_pos, self.last_gen = 0, 0
        self.chrom = chrom

    def _advance(self, pos, vcf_chrom=None):
        """Read map rows until one lies at or beyond pos (ref:192-211, 575-589)."""
        while pos > self.r_pos:
            self.last_r_pos, self.last_rate, self.last_cum = self.r_pos, self.rate, self.cum
            row = self.fh.readline()
            if row.strip() == '':                         # end of the map
                print(f'End-of-file for the recombination map. Assume uniform rate of {self.rec_rate} '
                      f'after {self.last_r_pos}.')
                self.r_pos, self.rate = 8e8, self.rec_rate * 1e6
                self.cum = self.last_cum + self.rec_rate * (8e8 - self.last_r_pos)
                break
            row = row.strip().split('\t')
            if vcf_chrom is not None:
                assert _INTS.findall(row[0])[0] == _INTS.findall(vcf_chrom)[0]
            elif row[0] not in {self.chrom, 'chr' + self.chrom}:
                print(row[0], self.chrom, row)
            self.r_pos, self.rate, self.cum = map(float, row[1:])

    def _interpolate(self, pos):
        span = (self.r_pos - self.last_r_pos)
        if self.last_pos < self.last_r_pos:
            return self.last_cum + (pos - self.last_r_pos) / span * (self.cum - self.last_cum)
        return self.last_gen + (pos - self.last_pos) / span * (self.cum - self.last_cum)

    def at(self, pos, vcf_chrom=None, anchor=None):
        """Genetic position of `pos`; the cursor remembers it.  `anchor`: the position the reference puts in the
        within-segment formula (it uses the VCF record's position for substitutions, ref:568)."""
        if pos <= self.r_pos:
            gen = self.last_gen + ((pos if anchor is None else anchor) - self.last_pos) * self.last_rate / 1e6
        else:
            self._advance(pos, vcf_chrom)
            gen = self._interpolate(pos)
        self.last_pos, self.last_gen = pos, gen
        return gen

    def close(self):
        self.fh.close()


def load_alignment(axtfile, chrom):
    """{position: (reference base, outgroup base)} and the sorted positions for one chromosome of a pairwise
    AXT alignment (ref:255-333)."""
    pairs, order, odd = {}, [], set()
    axt = _open_text(axtfile, '.axt')
    line = axt.readline()
    while line.startswith('#'):
        line = axt.readline()
    assert not re.match(r'^[A|T|C|G]', line)
    skipped = set()
    while line != '':
        head = line.strip().split(' ')
        if head == ['']:
            break
        primary = axt.readline().strip().upper()
        aligned = axt.readline().strip().upper()
        if head[1] in {str(chrom), 'chr' + str(chrom)}:
            pos = int(head[2]) - 1
            for i, base in enumerate(primary):
                if base in BASES:
                    pos += 1
                    if aligned[i] in BASES:
                        order.append(pos)
                        pairs[pos] = (base, aligned[i])
                else:
                    odd.add(base)
        else:
            skipped.add((head[1], head[4]))
        axt.readline()                                    # blank line
        line = axt.readline()                             # next summary line
    axt.close()
    order.sort()
    if skipped:
        print('skipped ', len(skipped), ' chromosomes:', sorted(skipped), 'chr' + chrom)
    print('non-SNP mappings:', odd)
    print(len(order), 'mapped positions recorded')
    assert len(order) == len(pairs)
    return pairs, order


def _polarise(ch, pos, pair, ref, alt):
    """True when the VCF's REF is the derived allele's complement, i.e. the count must be flipped (ref:437-452)."""
    a_ref, a_anc = pair
    if ref == a_ref and alt == a_anc:
        return False
    if ref == a_anc and alt == a_ref:
        return True
    if a_ref == a_anc and ref == a_ref:
        return True
    if a_ref == a_anc and alt == a_ref:
        return False
    print(f'On chr{ch} position {pos}, ref, anc in axt: ({a_ref}, {a_anc}); in vcf: ({ref}, {alt}).')
    sys.exit(1)


def _vcf_records(vcffile, pop_list):
    """-> (open file positioned after the #CHROM line, sample columns)."""
    vcf = _open_text(vcffile, '.vcf')
    line = vcf.readline()
    while line.startswith('##'):
        line = vcf.readline()
    return vcf, sample_columns(line, pop_list)


def parse_without_alignment(chrom, vcffile, rec_map, rec_rate, outfile, pop_list):
    """Folded counts (x = minor allele): for B0,MAF.  ref:56-128 (uniform rate) and ref:131-252 (map)."""
    out = open(outfile, 'w')
    out.write('position\tgenPos\tx\tn\n')
    vcf, cols = _vcf_records(vcffile, pop_list)
    cursor = MapCursor(rec_map, chrom, rec_rate) if rec_map is not None else None
    for line in vcf:
        f = line.strip().split('\t')
        if f[0] not in {chrom, 'chr' + chrom}:
            continue
        if cursor is not None and _INTS.findall(chrom)[0] != _INTS.findall(f[0])[0]:
            print('Please make sure the recombination map and vcf cover the same chromosome.')
            print(f[0])
            sys.exit(1)
        if not _is_counted_snp(f):
            continue
        if cursor is not None:
            pos = int(f[1])
            gen = cursor.at(pos)
        else:
            pos = f[1]
            gen = float(pos) * rec_rate
        x, n = count_alleles(f, cols)
        if 0 < x < n:
            out.write(f'{pos}\t{gen}\t{min(n - x, x)}\t{n}\n')
        # monomorphic among the chosen samples: the reference raises here, its shipped outputs skip (see above)
    vcf.close()
    out.close()
    if cursor is not None:
        cursor.close()


def parse_with_alignment(ch, axtfile, vcffile, rec_map, rec_rate, outfile, pop_list, ploidy):
    """Polarised counts plus substitutions: for B1, B2 and everything else.  ref:336-487 (uniform rate) and
    ref:493-706 (map)."""
    print(f'Loading the alignment for chr{ch}...')
    pairs, order = load_alignment(axtfile, ch)
    assert len(pairs) > 0 and len(order) > 0, order
    out = open(outfile, 'w')
    out.write('position\tgenPos\tx\tn\n')
    vcf, cols = _vcf_records(vcffile, pop_list)
    cursor = MapCursor(rec_map, ch, rec_rate) if rec_map is not None else None
    if cursor is None:
        print('Parsing ', len(cols), 'individuals of ', ploidy, 'ploidy.')
        vcf.readline()                                    # ref:361-362: the first record is read and dropped
    sample_size = ploidy * len(cols)
    k = 0                                                 # next aligned position not yet passed
    for line in vcf:
        f = line.strip().split('\t')
        if f[0] not in {ch, 'chr' + ch}:
            print(f[0], ch, {ch, 'chr' + ch})
            print(f[:11])
            sys.exit(1)
        pos = int(f[1])
        # substitutions between the previous record and this one
        while k < len(order) and order[k] < pos:
            a_pos = order[k]
            a_ref, a_anc = pairs[a_pos]
            if a_ref != a_anc:
                gen = a_pos * rec_rate if cursor is None else cursor.at(a_pos, f[0], anchor=pos)
                out.write(f'{a_pos}\t{gen}\t{sample_size}\t{sample_size}\n')
            k += 1
        if k == len(order):
            print('No more recorded positions.' if cursor is None else 'No more recorded positions in axt.')
            break
        if not _is_counted_snp(f):
            continue
        if pos not in pairs:
            print(f'{ch}, {pos}: not mapped')
            continue
        if len({pairs[pos][0], pairs[pos][1], f[3], f[4]}) > 2:
            print(f'{ch}, {pos}: not bi-allelic.\taxt-ref/anc: {pairs[pos]}; in vcf-ref/alt: {f[3]}/{f[4]}')
            continue
        assert order[k] == pos
        flip = _polarise(ch, pos, pairs[pos], f[3], f[4])
        x, n = count_alleles(f, cols)
        gen = pos * rec_rate if cursor is None else cursor.at(pos, f[0])
        if flip:
            if x < n:
                out.write(f'{pos}\t{gen}\t{n - x}\t{n}\n')
        elif x > 0:
            out.write(f'{pos}\t{gen}\t{x}\t{n}\n')
    vcf.close()
    out.close()
    if cursor is not None:
        cursor.close()


def build_parser():
    import argparse
    p = argparse.ArgumentParser(prog='parse_ballermix_input.py')
    p.add_argument('--vcf', dest='vcffile', required=True,
                   help='Path and name of the vcf file. Format can be either .vcf or .vcf.gz.\n')
    p.add_argument('-c', '--chr', dest='ch', required=True,
                   help='ID of the chromosome. E.g. 2a for chr2a, 12 for chr12, etc.\n')
    p.add_argument('-o', '--output', dest='outfile', required=True, help='Path and name of the output file.')
    p.add_argument('--ID_list', dest='pop_list', default=None,
                   help='Path and name to the file containing list of sample IDs (identical to their column names '
                        'in vcf) to be counted, separated by comma. If not provided, all samples in the vcf will be '
                        'counted.\n')
    p.add_argument('--axt', dest='axtfile', default=None,
                   help='Path and name of the sequence alignment (in .axt or .axt.gz) for calling substitution and '
                        'polarizing the allele frequency. If not provided, then the output will only be applicable '
                        'to B_0maf.')
    p.add_argument('--rec_rate', dest='rec_rate', type=float, default=1e-6,
                   help='Recombination rate in cM/nt. Default value is 1e-6 cM/nt.')
    p.add_argument('--rec_map', dest='rec_map', default=None,
                   help='Path and name of the recombination map (hapmap format) of the same sequence. If not '
                        'provided, a uniform recombination rate will be applied with a default rate of 1e-6 cM/nt. '
                        'Use "--rec_rate" to specify another rate.')
    p.add_argument('--hap', dest='hap', action='store_true', default=False,
                   help='Indicate that the organism in vcf is a haploid. Not necessary unless substitutions need to '
                        'be called.')
    return p


def main(argv=None):
    """Same dispatch and messages as ref:716-769."""
    argv = sys.argv[1:] if argv is None else list(argv)
    parser = build_parser()
    if len(argv) == 0:
        parser.print_help()
        sys.exit()
    opt = parser.parse_args(argv)
    last = time.time()
    ploidy = int(opt.hap) + (1 - opt.hap) * 2
    if opt.axtfile is not None:
        if opt.rec_map is not None:
            print(time.ctime(), 'Parsing BalLeRMix input for B_2 with genetic positions matching the recombination '
                  f'map. Assume genomic regions not covered by the map to recombine at {opt.rec_rate} cM/nt.')
        else:
            print(time.ctime(), f'Parsing BalLeRMix input for B_2 with a uniform recombination rate of '
                  f'{opt.rec_rate} cM/nt...')
        parse_with_alignment(opt.ch, opt.axtfile, opt.vcffile, opt.rec_map, opt.rec_rate, opt.outfile, opt.pop_list,
                             ploidy)
    else:
        if opt.rec_map is not None:
            print(time.ctime(), 'Parsing BalLeRMix input for B_0maf with genetic positions matching the '
                  f'recombination map. Assume genomic region not covered by the map to recombine at {opt.rec_rate} '
                  'cM/nt.')
        else:
            print(time.ctime(), f'Parsing BalLeRMix input for B_0maf with a uniform recombination rate of '
                  f'{opt.rec_rate} cM/nt...')
        parse_without_alignment(opt.ch, opt.vcffile, opt.rec_map, opt.rec_rate, opt.outfile, opt.pop_list)
    print(time.ctime(), f'Parsing completed. Output file: {opt.outfile} Total time {time.time() - last}sec.')


if __name__ == '__main__':
    main()
