"""
Generate fatiando_b200/csrc/prism_tables.h: the lookup tables and short polynomials of the
division-free logarithm and the table-reduced arctangent of prism_fastmath.cuh.

  log_tab(a),  a > 0 normal
      hi-word arithmetic splits a = 2^e * m, m in [M0, 2*M0), M0 = mk(LOG_OFF, 0) ~ 0.708, and
      picks bucket k = top 8 bits of (hi32(a) - LOG_OFF) mod 2^20.  c_k ~ 1/m_k:
          r = fma(m, c_k, -1)                         (one rounding, |r| <= RMAX)
          log(a) = e*ln2 - log(c_k) + (r + r^2*P(r))
      c_k is searched among the doubles next to 1/m_k for one whose -log(c_k) is a double to
      within 2^-10 ulp (Gal's accurate tables), so the table needs no low word; the bucket that
      holds m = 1 has c = 1, log(c) = 0 exactly (full RELATIVE accuracy for a near 1).

  atan_tab(a, b),  a, b >= 0, not both 0
      d = hi32(a) - hi32(b) is a piecewise-linear log2(a/b) in units of 2^-20 (off by
      +-0.0861); bucket k = clamp((d >> 16) + ATAN_BIAS, 0, ATAN_N - 1): 16 buckets per octave
      from 2^-6 to 2^6, k = 0 for smaller ratios (c = 0), k = ATAN_N-1 for larger ones.
          t = (a - c_k*b) / (b + c_k*a)               (two fma, one division, |t| <= TMAX)
          atan(a/b) = atan(c_k) + (t + t*u*Q(u)),  u = t*t
      c_k again searched so that atan(c_k) is a double to within 2^-10 ulp.

Run once, by hand (about a minute); the generated header is committed.
"""
import os
import struct
import sys
from multiprocessing import Pool

import mpmath as mp

mp.mp.prec = 140
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "fatiando_b200", "csrc", "prism_tables.h")

LOG_BITS = 8
LOG_N = 1 << LOG_BITS
# 1.0 (hi word 0x3ff00000) sits in the middle of a bucket: (0x3ff00000 - OFF) = 0x95800
LOG_OFF = 0x3ff00000 - 0x95800
ATAN_PER_OCT = 16
ATAN_OCT = 6
ATAN_N = 2 * ATAN_OCT * ATAN_PER_OCT + 2
ATAN_BIAS = ATAN_OCT * ATAN_PER_OCT + 1
PSEUDO_LOG_ERR = 0.0861          # max |log2(1+f) - f|
GAL_ULP = 2.0 ** -10


def mk(hi, lo=0):
    return struct.unpack('<d', struct.pack('<Q', ((hi & 0xffffffff) << 32) | (lo & 0xffffffff)))[0]


def next_double(x, k):
    u = struct.unpack('<q', struct.pack('<d', x))[0]
    return struct.unpack('<d', struct.pack('<q', u + k))[0]


def ulp(x):
    x = abs(x)
    return next_double(x, 1) - x


def gal_search(args):
    """first double c near `nominal` (stepping `stride` ulps) with f(c) within GAL_ULP ulp of a double"""
    kind, nominal = args
    f = (lambda c: -mp.log(c)) if kind == 'log' else mp.atan
    for k in range(200000):
        c = next_double(nominal, k * 4099)
        exact = f(mp.mpf(c))
        near = float(exact)
        if abs(exact - mp.mpf(near)) <= GAL_ULP * ulp(near):
            return c, near
    raise RuntimeError("no accurate table entry near %r" % nominal)


def cheb_fit(f, a, b, deg):
    n = deg + 1
    nodes = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for r, x in enumerate(nodes):
        for c in range(n):
            A[r, c] = x ** c
        y[r] = f(x)
    sol = mp.lu_solve(A, y)
    return [float(sol[i]) for i in range(n)]


def horner(coeffs, x):
    acc = mp.mpf(0)
    for c in reversed(coeffs):
        acc = acc * x + mp.mpf(c)
    return acc


def P_true(r):     # (log1p(r) - r) / r^2
    if abs(r) < mp.mpf(10) ** -12:     # the difference cancels: take the series
        return -mp.mpf(1) / 2 + r / 3 - r * r / 4
    return (mp.log1p(r) - r) / (r * r)


def Q_true(u):     # (atan(t)/t - 1)/u
    if u < mp.mpf(10) ** -24:
        return -mp.mpf(1) / 3 + u / 5
    t = mp.sqrt(u)
    return (mp.atan(t) / t - 1) / u


def main():
    # ---- logarithm ----
    log_nominal, rmax = [], 0.0
    for k in range(LOG_N):
        lo_hi = LOG_OFF + (k << (20 - LOG_BITS))
        hi_hi = LOG_OFF + ((k + 1) << (20 - LOG_BITS))
        m_lo, m_hi = mk(lo_hi), mk(hi_hi)
        if m_lo < 1.0 < m_hi:
            log_nominal.append(None)
            rmax = max(rmax, 1.0 - m_lo, m_hi - 1.0)
            continue
        c = 2.0 / (m_lo + m_hi)
        log_nominal.append(c)
        rmax = max(rmax, abs(m_lo * c - 1.0), abs(m_hi * c - 1.0))
    rmax *= 1.0 + 2.0 ** -30          # the Gal search moves c by < 2^-30 relative
    # ---- arctangent ----
    atan_nominal = [None]
    for k in range(1, ATAN_N - 1):
        atan_nominal.append(2.0 ** (-ATAN_OCT + (k - 1 + 0.5) / ATAN_PER_OCT))
    atan_nominal.append(2.0 ** (ATAN_OCT + 0.5))
    tmax = 0.0
    spread = 2.0 ** (0.5 / ATAN_PER_OCT + PSEUDO_LOG_ERR)
    for k, c in enumerate(atan_nominal):
        if c is None:               # ratios below 2^-OCT (+ pseudo-log error)
            tmax = max(tmax, 2.0 ** (-ATAN_OCT + PSEUDO_LOG_ERR))
        elif k == ATAN_N - 1:       # ratios above 2^OCT (- error) up to infinity
            lo = 2.0 ** (ATAN_OCT - PSEUDO_LOG_ERR)
            tmax = max(tmax, abs(lo - c) / (1 + lo * c), 1.0 / c)
        else:
            for rho in (c * spread, c / spread):
                tmax = max(tmax, abs(rho - c) / (1 + rho * c))
    tmax *= 1.0 + 2.0 ** -20
    jobs = [('log', c) for c in log_nominal if c is not None] + [('atan', c) for c in atan_nominal if c is not None]
    with Pool() as pool:
        found = pool.map(gal_search, jobs, chunksize=4)
    it = iter(found)
    log_tab = [(1.0, 0.0) if c is None else next(it) for c in log_nominal]
    atan_tab = [(0.0, 0.0) if c is None else next(it) for c in atan_nominal]

    tol = mp.mpf(2) ** -62
    for deg in range(2, 12):
        lg = cheb_fit(P_true, -mp.mpf(rmax), mp.mpf(rmax), deg)
        err = max(abs((x + x * x * horner(lg, x)) / mp.log1p(x) - 1)
                  for x in (mp.mpf(rmax) * (2 * mp.mpf(i) / 400 - 1) for i in range(401)) if x != 0)
        # error relative to log1p(r) itself matters only in the bucket of 1; elsewhere |log c| >> |r|
        if err < mp.mpf(2) ** -57:
            break
    lg_err = err
    umax = mp.mpf(tmax) ** 2
    for deg in range(2, 12):
        aq = cheb_fit(Q_true, mp.mpf(0), umax, deg)
        err = max(abs((1 + u * horner(aq, u)) / (1 + u * Q_true(u)) - 1)
                  for u in (umax * mp.mpf(i) / 400 for i in range(401)))
        if err < tol:
            break
    aq_err = err

    def arr(name, rows):
        body = ",\n    ".join("{%s, %s}" % (float.hex(a), float.hex(b)) for a, b in rows)
        return "static const double %s[%d][2] = {\n    %s};\n" % (name, len(rows), body)

    with open(OUT, "w") as fh:
        fh.write("// prism_tables.h -- GENERATED by tools/gen_tables.py; do not edit.\n")
        fh.write("// log_tab : %d buckets, |r| <= %.6g, P of degree %d, max rel. error of r + r^2 P(r) %.2e\n"
                 % (LOG_N, rmax, len(lg) - 1, float(lg_err)))
        fh.write("// atan_tab: %d buckets, |t| <= %.6g, Q of degree %d, max rel. error of t + t u Q(u) %.2e\n"
                 % (ATAN_N, tmax, len(aq) - 1, float(aq_err)))
        fh.write("// every table value f(c) is a double to within 2^-10 ulp (searched c)\n")
        fh.write("#pragma once\n#include \"prism_common.cuh\"\nnamespace fb200 {\nnamespace fm {\n")
        fh.write("constexpr int LOG_TAB_N = %d;\nconstexpr int LOG_TAB_SHIFT = %d;\n" % (LOG_N, 20 - LOG_BITS))
        fh.write("constexpr int LOG_TAB_OFF = 0x%08x;\n" % LOG_OFF)
        fh.write("constexpr int ATAN_TAB_N = %d;\nconstexpr int ATAN_TAB_BIAS = %d;\n" % (ATAN_N, ATAN_BIAS))
        fh.write("constexpr int ATAN_TAB_SHIFT = %d;\n" % (20 - 4))
        fh.write("constexpr int TAB_DOUBLES = 2 * (LOG_TAB_N + ATAN_TAB_N);\n")
        fh.write("// {c, -log(c)}\n" + arr("LOG_TAB_HOST", log_tab))
        fh.write("// {c, atan(c)}\n" + arr("ATAN_TAB_HOST", atan_tab))
        fh.write("#if defined(__CUDACC__)\n")
        fh.write("// one array in global memory, log table first: copied to shared memory by every block\n")
        both = log_tab + atan_tab
        fh.write("static __device__ const double TAB_DEV[TAB_DOUBLES] = {\n    %s};\n"
                 % ",\n    ".join("%s, %s" % (float.hex(a), float.hex(b)) for a, b in both))
        fh.write("static __constant__ double LG_C_DEV[%d] = {%s};\n" % (len(lg), ", ".join(float.hex(c) for c in lg)))
        fh.write("static __constant__ double AQ_C_DEV[%d] = {%s};\n" % (len(aq), ", ".join(float.hex(c) for c in aq)))
        fh.write("#endif\n")
        fh.write("static const double LG_C_HOST[%d] = {%s};\n" % (len(lg), ", ".join(float.hex(c) for c in lg)))
        fh.write("static const double AQ_C_HOST[%d] = {%s};\n" % (len(aq), ", ".join(float.hex(c) for c in aq)))
        for fname, name, coeffs in (("lg_poly", "LG_C", lg), ("aq_poly", "AQ_C", aq)):
            n = len(coeffs)
            fh.write("FB_HD double %s(double x)\n{\n" % fname)
            fh.write("#if defined(__CUDA_ARCH__)\n    const double *c = %s_DEV;\n#else\n    const double *c = %s_HOST;\n#endif\n" % (name, name))
            fh.write("    double p = c[%d];\n" % (n - 1))
            for k in range(n - 2, -1, -1):
                fh.write("    p = fma(p, x, c[%d]);\n" % k)
            fh.write("    return p;\n}\n")
        fh.write("}  // namespace fm\n}  // namespace fb200\n")
    print("log: rmax %.6g, degree %d, err %.2e; atan: tmax %.6g, degree %d, err %.2e"
          % (rmax, len(lg) - 1, float(lg_err), tmax, len(aq) - 1, float(aq_err)))


if __name__ == "__main__":
    sys.exit(main())
