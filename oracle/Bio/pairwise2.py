"""Stand-in for `Bio.pairwise2` -- TEST INFRASTRUCTURE ONLY, PARITY UNPINNED.

Biopython is a third-party dependency of the reference (README.md:19, version unpinned)
whose source is neither under /root/reference nor installed in this image.  The reference
touches it in exactly one place,

    pairwise2.align.globalms(fixed_seq, seq[good_before+1:good_after], 0, -1, -1, -1)[0]
                                                              (src/jasper.py:309)

and reads only `[0][0]` / `[0][1]` (the two gapped strings of the FIRST alignment that
Biopython's traceback completes).  Those strings only *describe* a multi-error patch as
fix.csv rows; the polished sequence never depends on them (src/jasper.py:312).

This module restates the published algorithm of Bio.pairwise2 (Biopython 1.7x,
`_make_score_matrix_fast` + `_recover_alignments` + `_find_gap_open`): affine-gap global
alignment with match 0 / mismatch -1 / gap-open -1 / gap-extend -1,
penalize_extend_when_opening=False (a gap of length n costs -n), end gaps penalised,
and the depth-first LIFO traceback whose per-cell preference order is
   1 gap-in-A open, 2 diagonal, 4 gap-in-B open, 8 gap-in-A extend, 16 gap-in-B extend,
with the "gap in A directly after a gap in B" dead-end rule.  It has NOT been checked
against a real Biopython (none is available offline); every consumer says so.
"""
from collections import namedtuple

Alignment = namedtuple("Alignment", "seqA seqB score start end")
MAX_ALIGNMENTS = 1000


def _gap_cost(length, open_, extend):
    # calc_affine_penalty(length, open, extend, penalize_extend_when_opening=False)
    if length <= 0:
        return 0.0
    return open_ + extend * length - extend


def _rint(x, precision=1000):
    return int(x * precision + 0.5)


def _score_matrices(a, b, match, mismatch, open_, extend):
    la, lb = len(a), len(b)
    first_gap = _gap_cost(1, open_, extend)
    score = [[None] * (lb + 1) for _ in range(la + 1)]
    trace = [[None] * (lb + 1) for _ in range(la + 1)]
    for i in range(la + 1):
        score[i][0] = _gap_cost(i, open_, extend)
    for j in range(lb + 1):
        score[0][j] = _gap_cost(j, open_, extend)
    col_score = [0] + [_gap_cost(j, 2 * open_, extend) for j in range(1, lb + 1)]
    for row in range(1, la + 1):
        row_score = _gap_cost(row, 2 * open_, extend)
        for col in range(1, lb + 1):
            nogap = score[row - 1][col - 1] + (match if a[row - 1] == b[col - 1] else mismatch)
            row_open = score[row][col - 1] + first_gap
            row_extend = row_score + extend
            row_score = max(row_open, row_extend)
            col_open = score[row - 1][col] + first_gap
            col_extend = col_score[col] + extend
            col_score[col] = max(col_open, col_extend)
            best = max(nogap, col_score[col], row_score)
            score[row][col] = best
            rs, cs = _rint(row_score), _rint(col_score[col])
            rt = (1 if _rint(row_open) == rs else 0) + (8 if _rint(row_extend) == rs else 0)
            ct = (4 if _rint(col_open) == cs else 0) + (16 if _rint(col_extend) == cs else 0)
            bs = _rint(best)
            t = 2 if _rint(nogap) == bs else 0
            if rs == bs:
                t += rt
            if cs == bs:
                t += ct
            trace[row][col] = t
    return score, trace


def _finish(a, b, ali_a, ali_b, row, col, gap):
    if row:
        ali_a += a[row - 1::-1]
    if col:
        ali_b += b[col - 1::-1]
    if row > col:
        ali_b += gap * (len(ali_a) - len(ali_b))
    elif col > row:
        ali_a += gap * (len(ali_b) - len(ali_a))
    return ali_a, ali_b


def _find_gap_open(a, b, ali_a, ali_b, row, col, col_gap, gap, score, trace, stack,
                   open_, extend, target, direction):
    dead_end = False
    target_score = score[row][col]
    for n in range(target):
        if direction == "col":
            col -= 1
            ali_a += gap
            ali_b += b[col:col + 1]
        else:
            row -= 1
            ali_a += a[row:row + 1]
            ali_b += gap
        actual = score[row][col] + _gap_cost(n + 1, open_, extend)
        if _rint(actual) == _rint(target_score) and n > 0:
            if not trace[row][col]:
                break
            stack.append((ali_a, ali_b, row, col, col_gap, trace[row][col]))
        if not trace[row][col]:
            dead_end = True
    return ali_a, ali_b, row, col, dead_end


def _recover(a, b, score, trace, open_, extend, gap="-", first_only=False):
    la, lb = len(a), len(b)
    best = score[la][lb]
    done = []
    stack = [("", "", la, lb, False, trace[la][lb])]
    while stack and len(done) < MAX_ALIGNMENTS:
        dead_end = False
        ali_a, ali_b, row, col, col_gap, t = stack.pop()
        while (row > 0 or col > 0) and not dead_end:
            cache = (ali_a, ali_b, row, col, col_gap)
            if not t:
                if col and col_gap:
                    dead_end = True
                else:
                    ali_a, ali_b = _finish(a, b, ali_a, ali_b, row, col, gap)
                break
            elif t % 2 == 1:            # open a gap in A (consume one base of B)
                t -= 1
                if col_gap:
                    dead_end = True
                else:
                    col -= 1
                    ali_a += gap
                    ali_b += b[col:col + 1]
                    col_gap = False
            elif t % 4 == 2:            # diagonal
                t -= 2
                row -= 1
                col -= 1
                ali_a += a[row:row + 1]
                ali_b += b[col:col + 1]
                col_gap = False
            elif t % 8 == 4:            # open a gap in B (consume one base of A)
                t -= 4
                row -= 1
                ali_a += a[row:row + 1]
                ali_b += gap
                col_gap = True
            elif t in (8, 24):          # extend a gap in A
                t -= 8
                if col_gap:
                    dead_end = True
                else:
                    col_gap = False
                    ali_a, ali_b, row, col, dead_end = _find_gap_open(
                        a, b, ali_a, ali_b, row, col, col_gap, gap, score, trace, stack,
                        open_, extend, col, "col")
            elif t == 16:               # extend a gap in B
                t -= 16
                col_gap = True
                ali_a, ali_b, row, col, dead_end = _find_gap_open(
                    a, b, ali_a, ali_b, row, col, col_gap, gap, score, trace, stack,
                    open_, extend, row, "row")
            if t:
                stack.append(cache + (t,))
            t = trace[row][col]
        if not dead_end:
            done.append((ali_a[::-1], ali_b[::-1], best, 0, None))
            if first_only:
                break
    out = []
    for al in done:
        if al not in out:
            out.append(al)
    return [Alignment(x[0], x[1], x[2], 0, len(x[0])) for x in out]


def first_global_alignment(a, b, match=0, mismatch=-1, open_=-1, extend=-1):
    """Only the first completed traceback -- what src/jasper.py:309 consumes."""
    if not a or not b:
        return []
    score, trace = _score_matrices(a, b, match, mismatch, open_, extend)
    return _recover(a, b, score, trace, open_, extend, first_only=True)


class _Align(object):
    def globalms(self, a, b, match, mismatch, open_, extend):
        if not a or not b:
            return []
        score, trace = _score_matrices(a, b, match, mismatch, open_, extend)
        return _recover(a, b, score, trace, open_, extend)


align = _Align()


def format_alignment(seqA, seqB, score, begin, end):
    return "%s\n%s\n  Score=%g\n" % (seqA, seqB, score)
