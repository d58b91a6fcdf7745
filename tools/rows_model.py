"""Lane-level Python model of the ROW-OWNED variant of kernel 2 (indelope_b200/csrc/ksw2_rows.cuh), for the unbanded call-site
(w < 0, zdrop < 0) of ksw_extz2_sse (src/ksw2/csrc/ksw2_extz2_sse.c:113-388).

It restates, one int8 lane at a time instead of four per 32-bit word, exactly what the CUDA code does differently from the SSE
code, so that those design decisions are checked on the CPU against the oracle's lane model (tests/test_rows_model.py):

  * a thread owns query ROWS (word w = rows 4w..4w+3 -> thread w % 8, slot w / 8); x, v stay, u, y move up one row per diagonal;
  * lanes in front of t = 0 or past t = tlen-1 are parked in a fixed state and kept out of the maximum; rows >= qlen carry a
    padding code that matches nothing;
  * the exact score H is carried along a row with u (H is a potential), as g = H + (q+e)(r+1) + bias;
  * every thread keeps the best score of its own rows, the first diagonal it was reached on and a snapshot of its scores there;
    the overall maximum and its position (SSE tie order :316-348) come out of the snapshots after the last diagonal;
  * mte / mqe / score are read off the rows that hold the cells in question;
  * the backtrack matrix is p[r][j] and is walked as ksw_backtrack (:47-79).

Test infrastructure: nothing in the product imports this file.
"""

TPAD, QPAD = 5, 6
NEG_INF = -0x40000000


def tie_rank(t, st0, en1):
    """order of the SSE arg-max over t in [st0, en0): four strided accumulators (lower accumulator, then lower t), then the scalar tail"""
    return 1 + ((((t - st0) & 3) if t < en1 else 4) << 20) + (t - st0)


def rows_align(query, target, match=1, mismatch=-2, q=5, e=1, W=5):
    """query/target: sequences of codes 0..4.  Returns (fields dict, cigar list of (op, len)) like oracle.pyoracle.ksw2."""
    qlen, tlen = len(query), len(target)
    out = dict(max=0, zdropped=0, max_q=-1, max_t=-1, mqe=NEG_INF, mqe_t=-1, mte=NEG_INF, mte_q=-1, score=NEG_INF, n_cigar=0,
               clamp_binds=0)  # clamp_binds: live cells where min(z, match + 2(q+e)) of :132 changed z (never, for real cells)
    if qlen <= 0 or tlen <= 0:
        return out, []
    assert qlen <= 32 * W
    qe, gbias = q + e, 2 * (q + e)
    maxsc = match + 2 * qe
    nrows = 32 * W
    tcode = lambda t: target[t] if 0 <= t < tlen else TPAD
    qcode = lambda j: query[j] if j < qlen else QPAD
    # boundary state of a row that has not reached t = 0 yet (and of a parked lane): x = 0, v = q, y = 0, u = 0
    x = [0] * nrows; v = [q] * nrows; u = [0] * nrows; y = [0] * nrows
    v[0] = 0                                     # v1 = 0 on diagonal 0 (:211)
    g = [q * (j - 1) - e + gbias for j in range(nrows)]  # H(-1, j) + (q+e) j + bias
    nr = qlen + tlen - 1
    p = [[0] * nrows for _ in range(nr)]
    owner = lambda j: (j >> 2) & 7               # thread of the group that owns row j
    tbest = [0] * 8; tr = [-1] * 8; snap = [None] * 8
    mte, mte_r, mqe, mqe_t, score = NEG_INF, -1, NEG_INF, -1, NEG_INF
    jl = qlen - 1
    for r in range(nr):
        uo, yo = u[:], y[:]
        goff = qe * (r + 1) + gbias
        dmax = [0] * 8
        for w in range(nrows // 4):
            lo = r - 4 * w                       # t of the first row of the word
            if not (0 <= lo <= tlen + 2 and 4 * w < qlen):
                continue                         # no lane of the word is inside the target
            for c in range(4):
                j = 4 * w + c; t = r - j
                ut = uo[j - 1] if j > 0 else (q if r else 0)   # :212 for row 0
                yt = yo[j - 1] if j > 0 else 0
                sc = match if tcode(t) == qcode(j) else mismatch
                if tcode(t) == 4 or qcode(j) == 4:
                    sc = 0
                for val in (x[j], v[j], ut, yt):
                    assert 0 <= val <= 63, "the carry-free form needs every byte in [0, 63]"
                z = sc + 2 * qe; a = x[j] + v[j]; b = yt + ut
                d = 1 if a > z else 0
                z = max(z, a)
                if b > z:
                    d = 2
                z = max(z, b)
                if z > maxsc and 0 <= t <= tlen - 1 and j < qlen:
                    out["clamp_binds"] += 1
                z = min(z, maxsc)
                un, vn = z - v[j], z - ut
                z -= q
                xn, yn = max(a - z, 0), max(b - z, 0)
                d |= (0x08 if a - z > 0 else 0) | (0x10 if b - z > 0 else 0)
                if 0 <= t <= tlen - 1:           # a live lane (rows >= qlen included: real cells of the query-extended problem)
                    x[j], v[j], u[j], y[j] = xn, vn, un, yn
                    p[r][j] = d
                    g[j] += un
                    dmax[owner(j)] = max(dmax[owner(j)], g[j])
                else:                            # parked
                    x[j], v[j], u[j], y[j] = 0, q, 0, 0
        for th in range(8):
            dm = dmax[th] - goff
            if dm > tbest[th]:
                tbest[th], tr[th], snap[th] = dm, r, g[:]
        if r >= tlen - 1:                        # en0 == tlen-1: H[en0] is the cell of row r - tlen + 1
            hen = g[r - tlen + 1] - goff
            if hen > mte:
                mte, mte_r = hen, r
        if r >= jl:                              # r - st0 == qlen-1: H[st0] is the cell of the last row
            h = g[jl] - goff
            if h > mqe:
                mqe, mqe_t = h, r - jl
            if r == nr - 1:
                score = h
    V = max(tbest)
    cands = [tr[th] for th in range(8) if tbest[th] == V and tr[th] >= 0]
    if V > 0 and cands:
        r = min(cands)
        st0, en0 = max(0, r - qlen + 1), min(r, tlen - 1)
        en1 = st0 + (((en0 - st0) >> 2) << 2)
        want = V + qe * (r + 1) + gbias
        best = None
        for th in range(8):
            if tbest[th] == V and tr[th] == r:
                for j in range(nrows):
                    t = r - j
                    if owner(j) == th and j < qlen and st0 <= t <= en0 and snap[th][j] == want:
                        rk = 0 if t == en0 else tie_rank(t, st0, en1)
                        best = rk if best is None or rk < best else best
        t = en0 if best == 0 else st0 + ((best - 1) & 0xfffff)
        out.update(max=V, max_t=t, max_q=r - t)
    out.update(mte=mte, mte_q=mte_r - ((tlen - 1) | 15), mqe=mqe, mqe_t=mqe_t, score=score)
    # ksw_backtrack from (tlen-1, qlen-1), is_rot = 1; unbanded, no state is ever forced
    i, j, state, cig = tlen - 1, qlen - 1, 0, []

    def push(op, n):
        if cig and cig[-1][0] == op:
            cig[-1][1] += n
        else:
            cig.append([op, n])
    while i >= 0 and j >= 0:
        tmp = p[i + j][j]
        if state == 0:
            state = tmp & 7
        elif not (tmp >> (state + 2)) & 1:
            state = 0
        if state == 0:
            state = tmp & 7
        if state == 0:
            push(0, 1); i -= 1; j -= 1
        elif state in (1, 3):
            push(2, 1); i -= 1
        else:
            push(1, 1); j -= 1
    if i >= 0:
        push(2, i + 1)
    if j >= 0:
        push(1, j + 1)
    cig.reverse()
    out["n_cigar"] = len(cig)
    return out, [(op, n) for op, n in cig]
