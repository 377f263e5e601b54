"""CPU restatement of JASPER's per-batch polishing loop -- TEST INFRASTRUCTURE ONLY.

This is the oracle for the path BASELINE.json's north_star names: src/jasper.py `main()`
(lines 12-32) and everything it calls.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product (jasper_b200/) never does.

Parity pin: tests/golden/*/ hold inputs and the byte-exact outputs of the UNMODIFIED
/root/reference/src/jasper.py run in the build container against the reference's own
Jellyfish binding (oracle/_ref); tests/test_oracle.py replays every golden case through
this file and compares fix rows, tallies and polished sequences.  One piece stays
PARITY-UNPINNED: the fix.csv *rows* of the multi-error (`nbad > k`) branch, which the
reference derives from Bio.pairwise2 (absent here; restated from its published algorithm in
oracle/Bio/pairwise2.py).  Polished sequences and tallies do not depend on it.  Likewise the
"Q value" TEXT of src/jasper.sh:239-257 (qv_like_bc below): GNU bc is absent here and its
scale rules and math-library series are restated from its published source; the integer
tallies it is computed from are pinned.

The restatement keeps the reference's observable semantics, including the ones that look
like accidents: Python slice clamping and negative-start wrap-around,
'A'-padding of short / invalid strings inside MerDNA, the sampled `check_sequence`,
banker's rounding, first-success enumeration orders, the un-rebased scan index after an
indel, and the crash paths (raised here as ReferenceCrash; the reference prints and exits 1).

`count` is any callable str -> int implementing C(s) = qf[jf.MerDNA(s).get_canonical()].
"""
from __future__ import annotations

from dataclasses import dataclass, field

DIVISOR = 50          # src/jasper.py:14
MIN_OVERLAP = 5       # src/jasper.py:532
SLACK = 10            # src/jasper.py:533
MAX_PATHS = 5000      # src/jasper.py:543
NO_FIX = "nN"         # src/jasper.py:228


class ReferenceCrash(Exception):
    """A path on which the reference raises inside its bare `except` and exits 1."""


@dataclass
class PassReport:
    rows: list = field(default_factory=list)   # [name, coord, fixed_base, original]  (src/jasper.py:219-222)
    wrong: int = 0                             # total_wrong_kmers (src/jasper.py:104)
    total: int = 0                             # total_kmers       (src/jasper.py:51)


@dataclass
class PolishResult:
    passes: list        # one PassReport per iteration 0..p (the last one is count-only)
    seqs: list          # [(name, str)] after the last fixing pass, in input order
    lookups: int = 0


def sampling_step(k: int) -> int:
    return max(2, round(k / 8))                # src/jasper.py:20


class PolishOracle:
    def __init__(self, count, k: int, solid_thre: int, align=None):
        self._count = count
        self.k = k
        self.solid = solid_thre
        self.step = sampling_step(k)
        self.lookups = 0
        self.branches = {}                     # branch name -> times taken (test coverage aid)
        if align is None:
            from Bio.pairwise2 import first_global_alignment as align  # oracle/Bio shim
        self._align = align

    def _hit(self, what: str):
        self.branches[what] = self.branches.get(what, 0) + 1

    # -- C(s) ---------------------------------------------------------------------------
    def C(self, s: str) -> int:
        self.lookups += 1
        return self._count(s)

    # -- src/jasper.py:585-599 ----------------------------------------------------------
    def passes_check(self, trial: str, thr) -> bool:
        k = self.k
        if self.C(trial[:k]) < thr:
            return False
        if self.C(trial[-k:]) < thr:
            return False
        for m in range(self.step, len(trial) - k, self.step):
            if self.C(trial[m:m + k]) < thr:
                return False
        return True

    def _weak_everywhere(self, s: str, thr) -> int:
        k = self.k
        return sum(1 for m in range(0, len(s) - k + 1) if self.C(s[m:m + k]) < thr)

    # -- src/jasper.py:392-431 ----------------------------------------------------------
    def try_substitution(self, tbf: str, thr):
        k = self.k
        bad = tbf[k - 1]
        for b in "ACTG":
            if b == bad:
                continue
            trial = tbf[:k - 1] + b + tbf[k:]
            if self.passes_check(trial, thr):
                return b, trial
        return None, None

    def try_drop_base(self, tbf: str, thr):
        k = self.k
        dropped = tbf[k - 1]
        trial = tbf[:k - 1] + tbf[k:]
        if self.passes_check(trial, thr):
            return dropped, trial
        return None, None

    def try_add_base(self, tbf: str, thr):
        k = self.k
        for b in "ATCG":
            trial = tbf[:k - 1] + b + tbf[k - 1:]
            if self.passes_check(trial, thr):
                return b, trial
        return None, None

    # -- src/jasper.py:340-388 ----------------------------------------------------------
    def try_two_ended(self, tbf: str, thr, full: str, gb: int, ga: int):
        k = self.k
        n = len(tbf)
        left_bad = tbf[n - k]
        right_bad = tbf[k - 1]
        g0 = max(0, gb - k + 1)
        na = int((k - 1 - n + k) / 2)
        if ga + k - 1 + na < len(full):
            after = full[ga + k - 1:ga + k - 1 + na]
        else:
            after = full[min(len(full) - 1, ga + k - 1):len(full)]
        bl = len(after)
        before = full[max(0, g0 - bl + 1):g0 + 1]
        for x in "ACTG":
            for y in "ACTG":
                if x == left_bad and y == right_bad:
                    continue
                if x != left_bad and y != right_bad:
                    continue
                trial = tbf[:n - k] + x + tbf[n - k + 1:k - 1] + y + tbf[k:]
                if self.passes_check(before + trial + after, thr):
                    if x == left_bad:
                        side = "e"
                    elif y == right_bad:
                        side = "s"
                    else:
                        continue
                    return x, y, side, trial
        return None, None, None, None

    # -- src/jasper.py:434-477 ----------------------------------------------------------
    def try_homopolymer_grow(self, tbf: str, thr):
        k = self.k
        if thr > self.solid:
            return None, None, None
        sb = tbf[k - 2]
        trial = tbf
        added = 0
        start_bad = len(tbf) - k + 1
        cur = start_bad
        while added < start_bad:
            trial = trial[:k - 1] + sb + trial[k - 1:]
            added += 1
            nb = self._weak_everywhere(trial, thr)
            if nb == 0:
                return k - 1, sb * added, trial
            if nb >= cur:
                break
            cur = nb
        for b in "ATCG":
            trial = tbf[:k - 2] + b + tbf[k - 2:]
            if self.passes_check(trial, thr):
                return k - 2, b, trial
        return None, None, None

    # -- src/jasper.py:479-524 ----------------------------------------------------------
    def try_homopolymer_shrink(self, tbf: str, thr):
        k = self.k
        if thr > self.solid:
            return None, None, None
        sb = tbf[k - 1]
        loc = tbf
        removed = 0
        start_bad = len(tbf) - k + 1
        cur = start_bad
        # the reference's loop guard tests the UNMODIFIED string, i.e. it is always true
        while removed < start_bad:
            cur -= 1
            removed += 1
            loc = loc[:k - 1] + loc[k:]
            if len(loc) == k:
                break
            nb = self._weak_everywhere(loc, thr)
            if nb == 0:
                return k - 1, sb * removed, loc
            if nb >= cur:
                break
            cur = nb
        for m in range(len(tbf) - k, len(tbf) - 1):
            trial = tbf[:m] + tbf[m + 1:]
            if self.passes_check(trial, thr):
                return m, tbf[m], trial
        return None, None, None

    # -- src/jasper.py:527-583 ----------------------------------------------------------
    def bridge(self, len_tbf: int, gkb: str, gka: str, thr):
        k = self.k
        if len(gkb) < k or len(gka) < k or thr > self.solid:
            return None
        alphabet = ["A", "C", "G", "T"]
        stem = gkb[0:k - 1]
        max_ext = round((len_tbf - 2 * k) * 1.2) + MIN_OVERLAP + SLACK
        min_patch = round((len_tbf - 2 * k) / 1.2) - SLACK
        paths = [gkb[k - 1:k]]
        for i in range(1, max_ext):
            paths = [p for p in paths if len(p) > 0]
            if len(paths) > MAX_PATHS:
                return None
            for p in range(len(paths)):
                if paths[p] == "":
                    continue
                km1 = (stem + paths[p])[-k + 1:]
                extended = 0
                for b in alphabet:
                    if self.C(km1 + b) >= thr:
                        tail = km1 + b
                        if i >= MIN_OVERLAP and i >= min_patch and tail[-MIN_OVERLAP:] == gka[0:MIN_OVERLAP]:
                            body = paths[p][:-1] if extended else paths[p]
                            joined = (stem + body + b + gka[-(k - MIN_OVERLAP):])[-(2 * k - 1):]
                            patch = (body + b)[1:-MIN_OVERLAP]
                            if self.passes_check(joined, thr):
                                return None if i == MIN_OVERLAP else patch
                        if not extended:
                            paths[p] += b
                            extended = 1
                        else:
                            paths.append(paths[p][:-1] + b)
                if not extended:
                    paths[p] = ""
        return None

    # -- src/jasper.py:226-332 ----------------------------------------------------------
    def repair(self, seq: str, tbf: str, thr, nbad: int, gb: int, ga: int):
        k = self.k
        s = max(0, gb - k + 2)
        fixed_base, original, idx = NO_FIX, "-", None

        def splice(sub):
            return seq[:s] + sub + seq[ga + k - 1:]

        def two_ended_or_shrink():
            nonlocal seq, fixed_base, original, idx
            left, right, side, sub = self.try_two_ended(tbf, thr, seq, gb, ga)
            if side is not None:
                self._hit("two_ended_" + side)
                if side == "s":
                    original, fixed_base, idx = "s" + seq[ga - 1], str(left), [ga - 1]
                else:
                    original, fixed_base, idx = "s" + seq[gb + 1], str(right), [gb + 1]
                seq = splice(sub)
                return
            at, base, sub = self.try_homopolymer_shrink(tbf, thr)
            if base is not None:
                self._hit("shrink_run" if at == k - 1 else "shrink_single")
                original, fixed_base, idx = "i" + base, "-", [at + s]
                seq = splice(sub)
            else:
                self._hit("unfixed_short")

        if nbad == k:
            b, sub = self.try_substitution(tbf, thr)
            if b is not None:
                self._hit("substitution")
                original, fixed_base, idx = "s" + seq[ga - 1], b, [ga - 1]
                seq = splice(sub)
            else:
                b, sub = self.try_drop_base(tbf, thr)
                self._hit("drop_base" if b is not None else "unfixed_k")
                if b is not None:
                    original, fixed_base, idx = "i" + seq[ga - 1], "-", [ga - 1]
                    seq = splice(sub)
        elif nbad == k - 1:
            b, sub = self.try_add_base(tbf, thr)
            if b is not None:
                self._hit("add_base")
                original, fixed_base, idx = "d-", b, [ga]
                seq = splice(sub)
            else:
                two_ended_or_shrink()
        elif 1 < nbad < k - 1 and len(tbf) >= k:
            at, base, sub = self.try_homopolymer_grow(tbf, thr)
            if base is not None:
                self._hit("grow_run" if at == k - 1 else "grow_single")
                original, fixed_base, idx = "d-", base, [at + s]
                seq = splice(sub)
            else:
                two_ended_or_shrink()
        elif nbad > k:
            gkb = seq[gb - k + 1:gb + 1]
            gka = seq[ga:ga + k]
            patch = self.bridge(len(tbf), gkb, gka, thr)
            self._hit("bridge" if patch is not None else "unfixed_long")
            if patch is not None:
                old = seq[gb + 1:ga]
                hits = self._align(patch, old)
                if not hits:
                    raise ReferenceCrash("pairwise2 returned no alignment (src/jasper.py:309)")
                new_rep, old_rep = hits[0][0], hits[0][1]
                seq = seq[:gb + 1] + patch + seq[ga:]
                idx, fixed_base, original = [], [], []
                for col in range(len(new_rep)):
                    o, c = old_rep[col], new_rep[col]
                    if c == o:
                        continue
                    if c == "-":
                        fixed_base.append("-"); original.append("i" + o)
                    elif o == "-":
                        fixed_base.append(c); original.append("d-")
                    else:
                        fixed_base.append(c); original.append("s" + o)
                    idx.append(col + gb + 1)
        return seq, fixed_base, original, idx

    # -- src/jasper.py:150-223 ----------------------------------------------------------
    def weak_run(self, i: int, seq: str, wrong: int, fix: bool, rows: list, name: str, rolling_thre):
        k, C = self.k, self.C
        thre = rolling_thre if rolling_thre > 0 else self.solid
        j = i - 1
        occ = C(seq[j:k + j])
        while occ < thre and j >= 0:
            j -= 1
            occ = C(seq[j:k + j])
        gb = j + k - 1
        prev = C(seq[j:k + j])
        cur = C(seq[i:k + i])
        if j == -1:
            gb = -1
        if rolling_thre == 0:
            while cur < thre and i < len(seq) - k + 1:
                i += 1
                cur = C(seq[i:k + i])
        else:
            while cur < thre and i < len(seq) - k + 1:
                if i - j > k:
                    return i + 1, seq, wrong, False
                i += 1
                cur = C(seq[i:k + i])
        ga = i
        if C(seq[gb - k + 2:gb + 2]) < self.solid / 2 and C(seq[gb - k + 3:gb + 3]) < self.solid / 2:
            pass
        elif rolling_thre == 0:
            while C(seq[gb - k + 2:gb + 2]) >= prev / 2 and gb - k + 1 < ga:
                if gb == -1:
                    break
                if prev >= thre / 2 and C(seq[gb - k + 2:gb + 2]) < thre / 2 and C(seq[gb - k + 3:gb + 3]) < thre / 2:
                    break
                prev = C(seq[gb - k + 2:gb + 2])
                gb += 1
            if gb >= len(seq) - 1:
                return i, seq, wrong, True
        s = max(0, gb - k + 2)
        if s + k + k >= len(seq):
            return s + k + k, seq, wrong, True
        if (C(seq[s + 1:s + k + 1]) < thre and C(seq[s + k - 2:s + k + k - 2]) < thre
                and C(seq[s + k - 1:s + k + k - 1]) < thre and C(seq[s + k:s + k + k]) >= thre):
            ga = s + k
        tbf = seq[s:ga + k - 1]
        nbad = len(range(s, ga))
        wrong += nbad
        if fix:
            if gb < 0:
                return i, seq, wrong, False
            seq, fixed_base, original, idx = self.repair(seq, tbf, thre, nbad, gb, ga)
            if fixed_base != NO_FIX:
                if len(idx) == 1:
                    rows.append([name, idx[0], fixed_base, original])
                else:
                    if len(idx) == 0:
                        raise ReferenceCrash("patch identical to original (src/jasper.py:221)")
                    rows.append([name, idx[0], fixed_base[0], original[0]])
                    rows.append([name, idx[1], fixed_base[1], original[1]])
        return i, seq, wrong, False

    # -- src/jasper.py:50-104 -----------------------------------------------------------
    def scan(self, name: str, seq: str, fix: bool, rows: list):
        k, C = self.k, self.C
        i = 0
        wrong = 0
        valid = frozenset("ACTGactg")
        while i < len(seq) - k + 1:
            w = seq[i:k + i]
            n = w.find("N")
            if n >= 0:
                i += n + 1
                continue
            n = w.find("n")
            if n >= 0:
                i += n + 1
                continue
            if not valid.issuperset(w):
                i += 1
                continue
            occ = C(w)
            if occ < self.solid:
                i, seq, wrong, stop = self.weak_run(i, seq, wrong, fix, rows, name, 0)
                if stop:
                    break
            elif i > 0 and occ < C(seq[max(0, i - k):max(k, i)]) / DIVISOR:
                total = 0
                ind = max(0, i - k)
                num = 0
                while ind < i:
                    num += 1
                    ind += self.step
                    total += C(seq[ind:k + ind])
                rt = round(total / num / DIVISOR)
                self._hit("rolling_weak" if occ < rt else "rolling_ok")
                if occ < rt:
                    i, seq, wrong, stop = self.weak_run(i, seq, wrong, fix, rows, name, round(total / num / 2))
                    if stop:
                        break
                else:
                    i += k - 1
            else:
                i += k - 1
        return seq, wrong

    # -- src/jasper.py:12-32, 35-130 ----------------------------------------------------
    def polish(self, seqs, passes: int, fix: bool = True) -> PolishResult:
        """seqs: [(name, str)] in file order (duplicate names collapse like the reference's dict)."""
        table = {}
        for name, s in seqs:
            table[name] = s
        reports = []
        for ite in range(passes + 1):
            do_fix = fix and ite != passes
            rep = PassReport()
            for name in list(table.keys()):
                seq = table[name]
                rep.total += len(seq) - self.k + 1
                seq, wrong = self.scan(name, seq, do_fix, rep.rows)
                table[name] = seq
                rep.wrong += wrong
            reports.append(rep)
        return PolishResult(passes=reports, seqs=list(table.items()), lookups=self.lookups)

    def branch_report(self) -> dict:
        return dict(sorted(self.branches.items()))


# ------------------------------------------------------------------------------- file formats
def read_fasta(path: str):
    """src/jasper.py:615-631: name = first token of the header minus '>', newlines stripped only."""
    seqs = {}
    name = None
    chunks = []
    with open(path, "r") as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    seqs[name] = "".join(chunks)
                name = line.split()[0][1:]
                chunks = []
            else:
                chunks.append(line.replace("\n", ""))
    if name is not None:
        seqs[name] = "".join(chunks)
    elif chunks:
        pass  # sequence before any header belongs to the discarded placeholder
    return list(seqs.items())


def format_fix_csv(rows) -> bytes:
    """src/jasper.py:114-119: csv.writer(delimiter=' '), CRLF line ends."""
    import csv
    import io
    buf = io.StringIO(newline="")
    w = csv.writer(buf, delimiter=" ")
    w.writerow(["Contig", "Base_coord", "Original", "Mutation"])
    w.writerows(rows)
    return buf.getvalue().encode()


def format_fasta(seqs, width: int = 60) -> bytes:
    """src/jasper.py:120-128, 142-147."""
    out = []
    for name, s in seqs:
        out.append(">%s\n" % name)
        for i in range(0, len(s), width):
            out.append(s[i:i + width] + "\n")
    return "".join(out).encode()


def format_tally(rep: PassReport) -> bytes:
    return ("%d %d\n" % (rep.wrong, rep.total)).encode()   # src/jasper.py:111


# ------------------------------------------------------------------------------- QV (bc emulation)
def qv_exact_math(wrong: int, total: int, k: int) -> str:
    """src/jasper.sh:239-246 with exact ln / exp truncated to bc's scales (the emulation of round 1): what bc would print if its
    l() and e() were exact.  Kept as a cross-check of qv_like_bc: the two may differ in the last digit only."""
    from decimal import Decimal, getcontext, ROUND_DOWN
    getcontext().prec = 120

    def trunc(x, scale):
        return x.quantize(Decimal(1).scaleb(-scale), rounding=ROUND_DOWN)

    if total == 0:
        raise ReferenceCrash("divide by zero in bc")
    pgood = Decimal(1) - trunc(Decimal(wrong) / Decimal(total), 10)
    if pgood <= 0:
        raise ReferenceCrash("bc: log of non-positive number")
    inv_k = trunc(Decimal(1) / Decimal(k), 50)
    ln_p = trunc(pgood.ln(), 50)
    err = Decimal(1) - trunc(trunc(ln_p * inv_k, 50).exp(), 50)
    if not err > 0:
        return "Inf"
    ln_e = trunc(err.ln(), 5)
    ln_10 = trunc(Decimal(10).ln(), 5)
    q = trunc(trunc(Decimal(-10) * ln_e, 5) / ln_10, 5)
    return _bc_text(q)


def _bc_text(q) -> str:
    s = format(q, "f")
    if s.startswith("0."):
        s = s[1:]
    elif s.startswith("-0."):
        s = "-" + s[2:]
    return s


class _BcDecimal:
    """GNU bc's arithmetic on `decimal` numbers (a number's scale = its exponent): every result truncated to the scale POSIX
    prescribes; l() and e() are the series of bc's own math library (libmath.b of GNU bc 1.06 / 1.07, restated from its
    published source; bc is absent here, so parity is unpinned -- DESIGN.md section 2)."""

    def __init__(self):
        from decimal import Decimal, Context, ROUND_DOWN
        self.D = Decimal
        self.ctx = Context(prec=400, rounding=ROUND_DOWN)

    def scale_of(self, x):
        return max(0, -x.as_tuple().exponent)

    def fix(self, x, scale):
        return x.quantize(self.D(1).scaleb(-scale), context=self.ctx)

    def add(self, a, b):
        return self.fix(self.ctx.add(a, b), max(self.scale_of(a), self.scale_of(b)))

    def sub(self, a, b):
        return self.fix(self.ctx.subtract(a, b), max(self.scale_of(a), self.scale_of(b)))

    def mul(self, a, b, scale):
        sa, sb = self.scale_of(a), self.scale_of(b)
        return self.fix(self.ctx.multiply(a, b), min(sa + sb, max(scale, sa, sb)))

    def div(self, a, b, scale):
        return self.fix(self.ctx.divide(a, b), scale)

    def sqrt(self, a, scale):
        # (exact square root truncated: context precision 400 digits is far beyond the 56 any caller keeps)
        return self.fix(self.ctx.sqrt(a), max(scale, self.scale_of(a)))

    def l(self, x, scale):            # libmath.b: define l(x)
        D = self.D
        if x <= 0:
            return self.div(self.sub(D(1), D(10) ** scale), D(1), scale)
        z, sc = scale, scale + 6
        f = D(2)
        while x >= 2:
            f = self.mul(f, D(2), sc)
            x = self.sqrt(x, sc)
        while x <= D("0.5"):
            f = self.mul(f, D(2), sc)
            x = self.sqrt(x, sc)
        n = self.div(self.sub(x, D(1)), self.add(x, D(1)), sc)
        v, m, i = n, self.mul(n, n, sc), 3
        while True:
            n = self.mul(n, m, sc)
            e = self.div(n, D(i), sc)
            if e == 0:
                return self.div(self.mul(f, v, sc), D(1), z)
            v = self.add(v, e)
            i += 2

    def e(self, x, scale):            # libmath.b: define e(x)
        D = self.D
        neg = x < 0
        if neg:
            x = -x
        z = scale
        n = self.add(D(6 + z), self.mul(D("0.44"), x, scale))
        sc, f = self.scale_of(x) + 1, 0
        while x > 1:
            f += 1
            x = self.div(x, D(2), sc)
            sc += 1
        sc = int(n)                    # `scale = n` keeps the integer part
        v, a, d, i = self.add(D(1), x), x, 1, 2
        while True:
            a = self.mul(a, x, sc)
            d *= i
            t = self.div(a, D(d), sc)
            if t == 0:
                for _ in range(f):
                    v = self.mul(v, v, sc)
                return self.div(D(1), v, z) if neg else self.div(v, D(1), z)
            v = self.add(v, t)
            i += 1


def qv_like_bc(wrong: int, total: int, k: int) -> str:
    """src/jasper.sh:239-246: the three bc invocations, each with bc's own scale rules and math-library series."""
    bc = _BcDecimal()
    D = bc.D
    if total == 0:
        raise ReferenceCrash("divide by zero in bc")
    pgood = bc.sub(D(1), bc.div(D(wrong), D(total), 10))                        # scale=10; 1-$wrong/$total
    if pgood <= 0:
        raise ReferenceCrash("bc: log of non-positive number")
    err = bc.sub(D(1), bc.e(bc.mul(bc.l(pgood, 50), bc.div(D(1), D(k), 50), 50), 50))   # scale=50; 1 - e(l($pgood)*(1/$KMER))
    if not err > 0:
        return "Inf"
    q = bc.div(bc.mul(D(-10), bc.l(err, 5), 5), bc.l(D(10), 5), 5)              # scale=5; -10*l($error_rate) / l(10)
    return _bc_text(q)
