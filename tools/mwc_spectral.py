"""Choose the multiplier of the 32-bit multiply-with-carry generator behind the ESIM noise streams.

x' = lo32(a*x + c), c' = hi32(a*x + c) is the linear congruential sequence s' = s / 2^32 mod m with m = a*2^32 - 1
(Marsaglia), so its t-dimensional structure is that of the lattice {(s, s*b, ..., s*b^(t-1)) / m}, b = 2^32.  For every
safe-prime modulus (period (m-1)/2 ~ 2^62) in a range the script computes the normalised spectral figures of merit
f_t = nu_t / (gamma_t^(1/2) m^(1/t)), t = 2..8 (Knuth 3.3.4), and prints the multipliers with the best worst-case figure.
"""
import itertools
import sys
from fractions import Fraction


def is_prime(n):
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def lll(B, delta=Fraction(99, 100)):
    n = len(B)
    B = [list(r) for r in B]

    def dot(u, v):
        return sum(x * y for x, y in zip(u, v))

    def gso():
        Bs, mu = [], [[Fraction(0)] * n for _ in range(n)]
        for i in range(n):
            v = [Fraction(x) for x in B[i]]
            for j in range(i):
                mu[i][j] = dot(B[i], Bs[j]) / dot(Bs[j], Bs[j])
                v = [a - mu[i][j] * b for a, b in zip(v, Bs[j])]
            Bs.append(v)
        return Bs, mu

    k = 1
    Bs, mu = gso()
    while k < n:
        for j in range(k - 1, -1, -1):
            q = round(mu[k][j])
            if q:
                B[k] = [a - q * b for a, b in zip(B[k], B[j])]
                Bs, mu = gso()
        if dot(Bs[k], Bs[k]) >= (delta - mu[k][k - 1] ** 2) * dot(Bs[k - 1], Bs[k - 1]):
            k += 1
        else:
            B[k], B[k - 1] = B[k - 1], B[k]
            Bs, mu = gso()
            k = max(k - 1, 1)
    return B


# Hermite constants gamma_t^t for t = 1..8
GAMMA_T = {2: Fraction(4, 3), 3: Fraction(2), 4: Fraction(4), 5: Fraction(8), 6: Fraction(64, 3), 7: Fraction(64), 8: Fraction(256)}


def figures(a, tmax=8, m=None, b=None):
    if m is None:
        m = a * 2 ** 32 - 1          # multiply-with-carry: modulus a*2^32-1, multiplier 2^32
        b = 2 ** 32
    out = []
    for t in range(2, tmax + 1):
        # dual lattice basis: (m,0,..), (-b,1,0,..), (-b^2,0,1,..) ...
        B = [[m] + [0] * (t - 1)]
        for i in range(1, t):
            r = [0] * t
            r[0] = -pow(b, i, m)
            r[i] = 1
            B.append(r)
        R = lll(B)
        best = min(sum(x * x for x in r) for r in R)
        rng = range(-2, 3) if t > 5 else range(-3, 4)
        for co in itertools.product(rng, repeat=t):
            if not any(co):
                continue
            v = [sum(c * R[i][j] for i, c in enumerate(co)) for j in range(t)]
            best = min(best, sum(x * x for x in v))
        nu2 = best
        # f_t = nu / (gamma_t^(1/2) * m^(1/t))
        f = (nu2 ** 0.5) / ((float(GAMMA_T[t]) ** (1.0 / t)) ** 0.5 * m ** (1.0 / t))
        out.append(f)
    return out


if __name__ == "__main__":
    lo, hi = (int(x, 0) for x in sys.argv[1:3]) if len(sys.argv) > 2 else (0xE0000000, 0xE0004000)
    res = []
    for a in range(lo, hi):
        m = a * 2 ** 32 - 1
        if is_prime(m) and is_prime((m - 1) // 2):
            f = figures(a)
            res.append((min(f), a, f))
    res.sort(reverse=True)
    for mn, a, f in res[:10]:
        print(hex(a), a, "min %.3f" % mn, " ".join("%.3f" % x for x in f))
