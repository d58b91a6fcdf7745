"""The design of the row-owned variant of kernel 2 (indelope_b200/csrc/ksw2_rows.cuh), checked on the CPU: tools/rows_model.py
restates it one lane at a time -- query rows owned by threads, lanes outside the target parked, exact scores carried along rows as
a potential, per-thread first-maximum snapshots, p[r][j] backtrack -- and must give the oracle's lane model of ksw_extz2_sse
(pinned to the reference's own compiled C file in test_oracle_ksw2.py) field for field and CIGAR for CIGAR."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from oracle import pyoracle as orc
from rows_model import rows_align

FIELDS = ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar")


def compare(qry, tgt, **kw):
    go = kw.get("q", 5); ge = kw.get("e", 1)
    fo, co, _ = orc.ksw2(np.asarray(qry, np.uint8), np.asarray(tgt, np.uint8), match=kw.get("match", 1), mismatch=kw.get("mismatch", -2), gapo=go, gape=ge, w=-1, zdrop=-1)
    fm, cm = rows_align([int(c) for c in qry], [int(c) for c in tgt], **kw)
    assert {k: fo[k] for k in FIELDS} == {k: fm[k] for k in FIELDS}, (len(qry), len(tgt))
    assert [(c & 0xf, c >> 4) for c in co] == cm, (len(qry), len(tgt))
    # H(t, j) - H(t-1, j-1) <= match for real cells: the clamp of ksw2_extz2_sse.c:132 never binds unbanded (the kernel keeps it anyway)
    assert fm["clamp_binds"] == 0


def test_known_answer():
    from test_oracle_ksw2 import QRY, TGT
    compare(orc.encode(QRY), orc.encode(TGT), q=3)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_pairs(seed):
    rng = np.random.default_rng(seed)
    for _ in range(120):
        ql = int(rng.integers(1, 70)); tl = int(rng.integers(1, 90))
        if rng.random() < 0.35:  # tandem repeats: many equal scores, the tie order decides
            unit = rng.integers(0, 4, int(rng.integers(1, 4)))
            q = np.resize(unit, ql).copy(); t = np.resize(unit, tl).copy()
        else:
            base = rng.integers(0, 4, ql + tl + 4); o = int(rng.integers(0, tl))
            q = base[o:o + ql].copy(); t = base[:tl].copy()
        for a in (q, t):
            m = rng.random(len(a)) < 0.04; a[m] = rng.integers(0, 5, int(m.sum()))  # substitutions and N
        compare(q, t)


def test_shapes_around_the_word_and_slot_boundaries():
    rng = np.random.default_rng(9)
    for ql in (1, 3, 4, 5, 31, 32, 33, 64, 65, 150, 160):
        for tl in (1, 2, 4, 5, 40):
            base = rng.integers(0, 4, ql + tl + 4)
            compare(base[min(2, tl - 1):min(2, tl - 1) + ql], base[:tl])


def test_other_scoring_parameters():
    rng = np.random.default_rng(4)
    for match, mismatch, q, e in ((2, -4, 4, 2), (1, -1, 2, 1), (3, -6, 10, 3)):
        for _ in range(25):
            ql = int(rng.integers(1, 50)); tl = int(rng.integers(1, 60)); base = rng.integers(0, 4, ql + tl + 4)
            o = int(rng.integers(0, tl))
            compare(base[o:o + ql], base[:tl], match=match, mismatch=mismatch, q=q, e=e)
