"""The fixed-point arithmetic of GNU bc, as far as src/jasper.sh:239-257 uses it.

jasper.sh turns the four QV tallies into the two "Q value" strings of its log with `bc` / `bc -l`:

    pgood      = "scale=10; 1-$wrong/$total"                 | bc
    error_rate = "scale=50; 1 - e(l($pgood)*(1/$KMER))"      | bc -l
    Q          = "scale=5; -10*l($error_rate) / l(10)"       | bc -l        (only if error_rate > 0, else "Inf")

bc is a third-party dependency of the reference (GNU bc, any 1.06/1.07: the math library has not changed between them) and is absent
from this image, so this module restates its published algorithm: numbers are decimal fixed point, every operation truncates
(never rounds) to a scale given by POSIX's rules, and `l()` / `e()` are the series of bc's own `libmath.b`, evaluated with exactly
those operations -- which is what decides the last printed digit.  Parity unpinned: no bc to run here.

  a + b, a - b   scale = max(scale(a), scale(b))
  a * b          scale = min(scale(a) + scale(b), max(scale, scale(a), scale(b)))
  a / b          scale = scale
  sqrt(a)        scale = max(scale, scale(a))

A number is (v, s): the integer v of units 10^-s.
"""
from math import isqrt


def _trunc_div(a: int, b: int) -> int:
    """a / b truncated toward zero (bc never rounds)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


class Bc:
    __slots__ = ("v", "s")

    def __init__(self, v: int, s: int = 0):
        self.v, self.s = v, s

    @staticmethod
    def parse(text: str) -> "Bc":
        """A bc literal / a number as bc printed it: the scale is the number of digits behind the point."""
        neg = text.startswith("-")
        t = text.lstrip("+-")
        ip, _, fp = t.partition(".")
        v = int((ip or "0") + fp)
        return Bc(-v if neg else v, len(fp))

    def text(self) -> str:
        """As bc prints it: all `scale` digits, no zero in front of the point."""
        a = abs(self.v)
        digits = str(a).rjust(self.s + 1, "0")
        ip, fp = (digits[:-self.s], digits[-self.s:]) if self.s else (digits, "")
        out = ("" if ip == "0" and fp else ip) + ("." + fp if fp else "")
        return ("-" if self.v < 0 else "") + out

    def at(self, s: int) -> int:
        """The value in units of 10^-s (s >= self.s)."""
        return self.v * 10 ** (s - self.s)

    def is_zero(self) -> bool:
        return self.v == 0


def add(a: Bc, b: Bc) -> Bc:
    s = max(a.s, b.s)
    return Bc(a.at(s) + b.at(s), s)


def sub(a: Bc, b: Bc) -> Bc:
    s = max(a.s, b.s)
    return Bc(a.at(s) - b.at(s), s)


def mul(a: Bc, b: Bc, scale: int) -> Bc:
    s = min(a.s + b.s, max(scale, a.s, b.s))
    return Bc(_trunc_div(a.v * b.v, 10 ** (a.s + b.s - s)), s)


def div(a: Bc, b: Bc, scale: int) -> Bc:
    if b.v == 0:
        raise ZeroDivisionError("bc: divide by zero")
    # (a.v / 10^a.s) / (b.v / 10^b.s) in units of 10^-scale
    num, den = a.v * 10 ** (scale + b.s), b.v * 10 ** a.s
    return Bc(_trunc_div(num, den), scale)


def sqrt(a: Bc, scale: int) -> Bc:
    s = max(scale, a.s)
    return Bc(isqrt(a.v * 10 ** (2 * s - a.s)), s)


def cmp(a: Bc, b: Bc) -> int:
    s = max(a.s, b.s)
    x, y = a.at(s), b.at(s)
    return (x > y) - (x < y)


ONE = Bc(1)
TWO = Bc(2)
HALF = Bc(5, 1)


def l(x: Bc, scale: int) -> Bc:
    """libmath.b l(x): ln x = 2 (a + a^3/3 + a^5/5 + ...), a = (x-1)/(x+1), after square roots have brought x into (.5, 2);
    summed at scale + 6 until a term truncates to zero."""
    if x.v <= 0:
        return div(sub(ONE, Bc(10 ** scale)), ONE, scale)
    z, sc = scale, scale + 6
    f = Bc(2)
    while cmp(x, TWO) >= 0:
        f = mul(f, TWO, sc)
        x = sqrt(x, sc)
    while cmp(x, HALF) <= 0:
        f = mul(f, TWO, sc)
        x = sqrt(x, sc)
    n = div(sub(x, ONE), add(x, ONE), sc)
    v = n
    m = mul(n, n, sc)
    i = 3
    while True:
        n = mul(n, m, sc)
        e = div(n, Bc(i), sc)
        if e.is_zero():
            return div(mul(f, v, sc), ONE, z)
        v = add(v, e)
        i += 2


def e(x: Bc, scale: int) -> Bc:
    """libmath.b e(x): e^x = (e^(x/2^f))^(2^f), the power series summed at scale 6 + scale + .44 x until a term truncates to zero."""
    neg = x.v < 0
    if neg:
        x = Bc(-x.v, x.s)
    z = scale
    n = add(Bc(6 + z), mul(Bc(44, 2), x, scale))        # n = 6 + z + .44*x; `scale = n` keeps the integer part
    sc = x.s + 1
    f = 0
    while cmp(x, ONE) > 0:
        f += 1
        x = div(x, TWO, sc)
        sc += 1
    sc = _trunc_div(n.v, 10 ** n.s)
    v = add(ONE, x)
    a = x
    d = 1
    i = 2
    while True:
        a = mul(a, x, sc)
        d *= i
        t = div(a, Bc(d), sc)
        if t.is_zero():
            for _ in range(f):
                v = mul(v, v, sc)
            return div(ONE, v, z) if neg else div(v, ONE, z)
        v = add(v, t)
        i += 1


def q_value(wrong: int, total: int, k: int) -> str:
    """The "Q value" text of src/jasper.sh:239-246 / :249-256 for one pair of tallies."""
    pgood = sub(ONE, div(Bc(wrong), Bc(total), 10))                                   # scale=10; 1-$wrong/$total
    pgood = Bc.parse(pgood.text())                                                    # (through the shell variable)
    rate = sub(ONE, e(mul(l(pgood, 50), div(ONE, Bc(k), 50), 50), 50))                # scale=50; 1 - e(l($pgood)*(1/$KMER))
    rate = Bc.parse(rate.text())
    if not cmp(rate, Bc(0)) > 0:                                                      # (( $(echo "$error_rate > 0" | bc -l) ))
        return "Inf"
    q = div(mul(Bc(-10), l(rate, 5), 5), l(Bc(10), 5), 5)                             # scale=5; -10*l($error_rate) / l(10)
    return q.text()
