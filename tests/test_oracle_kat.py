"""The reference's own in-file known-answer tests, run against the CPU oracle.

Vectors: src/contig.nim:293-430 (suite "contig"), src/genotyper.nim:49-70, src/indelope.nim:23-38.
These pin the oracle's reading of the assembler semantics; the GPU path is then compared to the
oracle in tests/test_gpu_*.py.
"""
import math

import pytest

from oracle import pyoracle as orc

UNALIGNED = -(2 ** 63)
T15 = "TTAACTGGGTACGGT"


def test_slide_align_offsets():  # src/contig.nim:293-321
    sa = orc.slide_align("ACTGGGTACGGT", T15, min_overlap=5)
    assert (sa.offset, sa.matches) == (3, 12)
    assert orc.slide_align("ACTGGGTACGGTGGG", T15, min_overlap=5).offset == 3
    assert orc.slide_align("ACTGGGTACG", T15, min_overlap=5).offset == 3
    assert orc.slide_align(T15, T15, min_overlap=5).offset == 0
    assert orc.slide_align("ATTAACTGGGTACGGT", T15, min_overlap=5).offset == -1
    assert orc.slide_align("ATTAACTGGGTACGGT", "TTAACTGGGTACGGTTTT", min_overlap=5).offset == -1
    assert orc.slide_align("ATTAACTGGGTACGGTTTGGGG", "TTAACTGGGTACGGTTTG", min_overlap=5).offset == -1
    m = orc.slide_align("ATTAACTGGGTACGGTTTGGGG", "TTAACTGGGTACGGTTTG", min_overlap=50)
    assert m.offset == UNALIGNED and not m.aligned


def test_corrections():  # src/contig.nim:323-343
    t, q = "ATTAACTGGGTACGGTTTGGGG", "TTAACTGGGXACGGTTTGG"
    ma = orc.slide_align(q, t, min_overlap=5, qsup=6, tsup=2, rule=1)
    assert ma.n_corr == 0
    ma = orc.slide_align(q, t, min_overlap=5, qsup=7, tsup=2, rule=1)
    assert ma.n_corr == 1
    assert q[ma.corr[0]] == "X" and t[ma.corr[1]] == "T" and ma.corr[2] == 1
    t, q = "ATTAACTGGGAACGGTTTGGGG", "GGAGATTAACTGGGXACGGTTTGG"
    ma = orc.slide_align(q, t, min_overlap=5, qsup=2, tsup=7, rule=1)
    assert ma.n_corr == 1
    assert q[ma.corr[0]] == "X" and t[ma.corr[1]] == "A" and ma.corr[2] == 0


def test_insertion_left_overhang():  # src/contig.nim:356-389
    t, q = "ATTAACTGGGTACGGTTTGGGG", "GGAGATTAACTGGGXACGGTTTGG"
    ma = orc.slide_align(q, t, min_overlap=5, qsup=2, tsup=7, rule=1)
    assert ma.aligned
    seq, sup, start, _ = orc.insert(t, 3, 7, q, 1, 2, ma)
    assert seq == "GGAGATTAACTGGGTACGGTTTGGGG" and len(sup) == 26 and start == 1
    assert sup == [2, 2, 2, 2, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 7, 9, 9, 9, 9, 9, 9, 9, 9, 9, 7, 7]

    ma = orc.slide_align(q, t, min_overlap=5, qsup=7, tsup=2, rule=1)
    seq, sup, start, _ = orc.insert(t, 5, 2, q, 0, 7, ma)
    assert start == 0 and ma.aligned
    assert seq == "GGAGATTAACTGGGXACGGTTTGGGG"
    assert sup == [7, 7, 7, 7, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 7, 9, 9, 9, 9, 9, 9, 9, 9, 9, 2, 2]

    t = "ATTAACTGGGTAC"
    ma = orc.slide_align(q, t, min_overlap=5, qsup=2, tsup=7, rule=1)
    assert ma.aligned
    seq, sup, start, _ = orc.insert(t, 3, 7, q, 0, 2, ma)
    assert seq == "GGAGATTAACTGGGTACGGTTTGG"
    assert sup == [2, 2, 2, 2, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 7, 9, 9, 2, 2, 2, 2, 2, 2, 2]
    assert start == 0


def test_insertion_right_overhang():  # src/contig.nim:391-422
    t, q = "GGAGATTAACTGGGXACGGTTTGG", "ATTAACTGGGTACGGTTTGGGG"
    ma = orc.slide_align(q, t, min_overlap=5, qsup=7, tsup=2, rule=1)
    assert ma.aligned
    seq, sup, start, _ = orc.insert(t, 1, 2, q, 3, 7, ma)
    assert start == 1
    assert sup == [2, 2, 2, 2, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 7, 9, 9, 9, 9, 9, 9, 9, 9, 9, 7, 7]
    assert seq == "GGAGATTAACTGGGTACGGTTTGGGG"

    q = "GGAGATTAACTGGGTACGGTTTGGGG"
    ma = orc.slide_align(q, t, min_overlap=5, qsup=2, tsup=7, rule=1)
    assert ma.offset == 0 and ma.aligned
    seq, sup, start, _ = orc.insert(t, 90, 7, q, 90, 2, ma)
    assert start == 90 and len(seq) == 26 and seq == "GGAGATTAACTGGGXACGGTTTGGGG"

    q = "AAAGGAGATTAACTGGGTACGGTTTGGGG"
    ma = orc.slide_align(q, t, min_overlap=5, qsup=7, tsup=2, rule=1)
    assert ma.offset == -3
    seq, sup, start, _ = orc.insert(t, 0, 2, q, 3, 7, ma)
    assert len(seq) == len(q) and seq == q and start == 3
    assert sup == [7, 7, 7, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 7, 9, 9, 9, 9, 9, 9, 9, 9, 9, 7, 7]


def test_insert_contained():  # src/contig.nim:424-430
    m = orc.Match()
    m.matches, m.offset, m.mismatches, m.aligned, m.n_corr = 19, 3, 0, 1, 0
    seq, sup, _, _ = orc.insert("CCGGGCTGGGCTT", 1, 2, "GGCTGGGCT", 1, 2, m)
    assert sup == [2, 2, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 2]


def test_first_accept_threshold_is_min_overlap_minus_one():
    # src/contig.nim:81-82,107: best_ma starts at min_overlap-1 and best_mm at max_mismatch+1, so the FIRST candidate
    # is accepted with ma == min_overlap-1 (mm 0 < 1). A literal reading, not covered by the reference's tests.
    m = orc.slide_align("ACGTA", "ACGTAC", min_overlap=6)
    assert m.aligned and m.matches == 5 and m.offset == 0
    assert not orc.slide_align("ACGTA", "ACGTACG", min_overlap=7).aligned


def test_genotyper():  # src/genotyper.nim:49-70
    HOM_REF, HET, HOM_ALT, UNKNOWN = 0, 1, 2, 3
    assert orc.genotype(10, 10, 1e-4)[0] == HET
    assert orc.genotype(20, 0, 1e-4)[0] == HOM_REF
    assert orc.genotype(1, 19, 1e-2)[0] == HOM_ALT
    assert orc.genotype(1, 19, 1e-8)[0] == HET
    assert orc.genotype(0, 0, 1e-8)[0] == UNKNOWN
    g, text, _ = orc.genotype(1, 19, 1e-8)
    assert text.startswith("0/1:")
    # renderings recomputed with C sprintf("%#.4f")
    assert orc.genotype(10, 10, 1e-4)[1] == "0/1:78.2415:-92.1044,-13.8629,-92.1044"
    assert text == "0/1:4.5577:-349.9929,-13.8629,-18.4207"
    # closed form check of eqn 2 for the het class
    _, _, q = orc.genotype(10, 10, 1e-4)
    gl1 = -20 * math.log(2) + 20 * math.log(1.0)
    gl0 = -20 * math.log(2) + 10 * math.log(2 * (1 - 1e-4)) + 10 * math.log(2 * 1e-4)
    assert abs(q - (gl1 - gl0)) < 1e-9


def test_read_trim():  # src/indelope.nim:23-38
    assert orc.trim([30] * 10) == (0, 10)
    assert orc.trim([2, 2, 30, 30, 30, 2]) == (2, 3)
    assert orc.trim([2] * 9 + [30]) == (9, 0)   # left scan stops at high: the read becomes empty
    assert orc.trim([2] * 10) == (9, 0)
    assert orc.trim([30]) == (0, 0)             # a == high == 0
    assert orc.trim([30, 2]) == (0, 1)


def test_mincode_is_canonical():
    a = "ACGTTGCAAGGCTTAACCGGTTAAGCA"
    rc = a[::-1].translate(str.maketrans("ACGT", "TGCA"))
    assert orc.mincode(a) == orc.mincode(rc)
    assert orc.mincode(a) != orc.mincode("C" + a[1:])
    assert orc.mincode("N" + a[1:]) is None


@pytest.mark.parametrize("name", ["pr1_small", "tandem_small"])
def test_oracle_reproduces_the_golden_fixtures(name):
    """tests/golden/* were made by tools/make_golden.py with the oracle calling the reference's own compiled ksw2_extz2_sse.c
    (oracle/_ref); the oracle with its own lane-exact DP, which is all that exists on a box without /root/reference, must give
    the same bytes"""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    from make_golden import FIXTURES
    from indelope_b200 import host
    cfg = dict(host.CONFIGS["pr1"]); cfg.update(FIXTURES[name])
    rois = host.Dataset(**cfg).sweep(min_reads=5)
    dump, vcf, cnt = orc.call(rois.arrays(), min_reads=5, min_ctg_len=73, min_event_len=5, use_ref_ksw2=False, dump_level=31)
    gold = os.path.join(root, "tests", "golden", name)
    assert rois.header() + vcf == open(gold + ".vcf").read()
    assert "\n".join(l for l in dump.splitlines() if l[:1] in "RAEV") + "\n" == open(gold + ".dump").read()
    if name == "tandem_small":
        assert cnt["al_events"] >= 20 and cnt["dp_b"] > 1000
