"""Prototype (host, numpy) of the domain-restricted x**kappa used by the Exner function:
exp2(kappa * log2(x)) with a lean log2 (atanh series) and exp2 (Taylor), to size the
polynomial degrees against a 50-digit reference before writing the device version."""
import numpy as np
from decimal import Decimal, getcontext

getcontext().prec = 50
LD = np.longdouble


def fma(a, b, c):
    return np.float64(LD(a) * LD(b) + LD(c))


LN2 = Decimal(2).ln()
# log(m) = 2 s + s z P(z),  z = s^2,  P(z) = 2/3 + 2/5 z + 2/7 z^2 + ...
NLOG = 11
CLOG = [np.float64(Decimal(2) / Decimal(2 * k + 3)) for k in range(NLOG)]
L2E = Decimal(1) / LN2
L2E_HI = np.float64(L2E)
L2E_LO = np.float64(L2E - Decimal(float(L2E_HI)))
NEXP = 14
CEXP = [np.float64(LN2 ** k / Decimal(np.math.factorial(k) if hasattr(np, "math") else __import__("math").factorial(k))) for k in range(NEXP + 1)]


def log2_lean(x):
    x = np.asarray(x, dtype=np.float64)
    bits = x.view(np.int64)
    e = ((bits >> 52) & 0x7FF) - 1023
    mb = (bits & 0x000FFFFFFFFFFFFF) | (1023 << 52)
    m = mb.view(np.float64)
    big = m > np.float64(1.4142135623730951)
    m = np.where(big, m * 0.5, m)
    e = e + big
    f = m - 1.0
    d = m + 1.0
    r = 1.0 / d                       # device: rcp approximation + Newton; here correctly rounded
    s = f * r
    s = fma(fma(-d, s, f), r, s)      # one correction step
    z = s * s
    p = CLOG[-1]
    for c in CLOG[-2::-1]:
        p = fma(p, z, c)
    # log(m) = 2 s + s z p, kept as hi + lo
    t = s * z
    lo = t * p
    hi = 2.0 * s
    lm_hi = hi + lo
    lm_lo = (hi - lm_hi) + lo
    # log2(x) = e + lm * log2(e)
    a = lm_hi * L2E_HI
    a_err = fma(lm_hi, L2E_HI, -a)
    b = fma(lm_hi, L2E_LO, fma(lm_lo, L2E_HI, a_err))
    return e.astype(np.float64), a, b   # log2 = e + a + b  (unevaluated)


def pow_lean(x, kappa):
    e, a, b = log2_lean(x)
    # y = kappa * (e + a + b), split as n + r with the large parts cancelled exactly
    ye = kappa * e
    ye_err = fma(kappa, e, -ye)
    ya = kappa * a
    ya_err = fma(kappa, a, -ya)
    n = np.rint(ye + ya)
    r = ((ye - n) + ya) + (ye_err + (ya_err + kappa * b))
    p = CEXP[-1]
    for c in CEXP[-2::-1]:
        p = fma(p, r, c)
    return np.ldexp(p, n.astype(np.int64))


def ulp_err(got, x, kappa):
    worst = 0.0
    for g, xv in zip(got, x):
        ref = (Decimal(float(xv)).ln() * Decimal(float(kappa))).exp()
        err = abs(Decimal(float(g)) - ref) / ref
        worst = max(worst, float(err) / 2.0 ** -53)
    return worst


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    kappa = np.float64(287.05 / 1004.0)
    for lo, hi in ((0.005, 0.1), (0.1, 0.6), (0.6, 1.3), (1.3, 4.0)):
        x = rng.uniform(lo, hi, size=4000)
        got = pow_lean(x, kappa)
        ref = np.power(x, kappa)
        print(f"x in [{lo}, {hi}]: lean {ulp_err(got, x, kappa):.3f} ulp, numpy pow {ulp_err(ref, x, kappa):.3f} ulp,"
              f" max |lean-numpy| {np.max(np.abs(got - ref) / ref) / 2.0 ** -53:.3f} ulp")
