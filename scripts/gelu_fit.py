"""Derivation of the one-MUFU exact-GELU used by the tile-plan decoder tail (csrc/decoder_tail_plan.cuh, TPG_COEFFS).

    erfc(z) = exp(-z^2) erfcx(z);  erfcx(a / sqrt 2) ~ Q(a), degree-10 polynomial in a = |x|, fitted in the weighted
    minimax sense (Lawson iteration, weight exp(-z^2)) so that the error of erfc itself is uniformly small.
    gelu(x) = max(x, 0) - (a Q(a) / 2) exp(-x^2 / 2),   gelu'(x) = Phi(x) + x exp(-x^2/2) / sqrt(2 pi).

Prints the coefficients -Q_k/2 (as committed) and the fp32-emulated error of gelu / gelu' against fp64 erf, next to the
Abramowitz-Stegun 7.1.26 form used by the other tail kernels.  Needs scipy; run on the CPU.
"""
import numpy as np
import numpy.polynomial.chebyshev as C
import numpy.polynomial.polynomial as Pn
from scipy.special import erf, erfcx

DEG, ZMAX = 10, 4.0


def fit(n=DEG, zmax=ZMAX, iters=400):
    z = np.linspace(0, zmax, 6001)
    f, wgt = erfcx(z), np.exp(-z * z)
    V = C.chebvander(2 * z / zmax - 1, n)
    lw = np.ones_like(z)
    for _ in range(iters):
        sw = np.sqrt(lw) * wgt
        c, *_ = np.linalg.lstsq(V * sw[:, None], f * sw, rcond=None)
        err = np.abs((V @ c - f) * wgt)
        lw = lw * err
        lw /= lw.sum()
    cu = C.cheb2poly(c)
    pz = np.zeros(1)
    for k, ck in enumerate(cu):
        pz = Pn.polyadd(pz, ck * Pn.polypow([-1.0, 2 / zmax], k))
    return pz / (np.sqrt(2.0) ** np.arange(len(pz))), err.max()      # polynomial in a = sqrt(2) z


def main():
    q, werr = fit()
    print("weighted max error of erfc: %.3e" % werr)
    print("TPG_COEFFS = {" + ", ".join("%.11gf" % (-0.5 * v) for v in q) + "}")
    f32 = np.float32
    x = np.linspace(-10, 10, 400001).astype(f32)
    a, xd = np.abs(x), x.astype(np.float64)
    u = f32(-0.5 * q[-1]) * np.ones_like(a)
    for ck in q[-2::-1]:
        u = (u * a + f32(-0.5 * ck)).astype(f32)
    e = np.exp2(((x * x).astype(f32) * f32(-0.72134752044448170368)).astype(f32)).astype(f32)
    u = (u * e).astype(f32)
    g = (u * a + np.maximum(x, 0)).astype(f32)
    phi = (np.copysign((u + f32(0.5)).astype(f32), x) + f32(0.5)).astype(f32)
    dg = ((x * f32(0.3989422804014327)).astype(f32) * e + phi).astype(f32)
    gt = xd * 0.5 * (1 + erf(xd / np.sqrt(2)))
    dgt = 0.5 * (1 + erf(xd / np.sqrt(2))) + xd * np.exp(-xd * xd / 2) / np.sqrt(2 * np.pi)
    print("polynomial: max |gelu err| / max(1,|x|) = %.3e, max |gelu' err| = %.3e"
          % ((np.abs(g - gt) / np.maximum(1, np.abs(xd))).max(), np.abs(dg - dgt).max()))
    z = (a * f32(0.70710678)).astype(f32)
    t = (f32(1) / (f32(0.3275911) * z + f32(1))).astype(f32)
    ee = np.exp(-(z * z)).astype(f32)
    poly = f32(1.061405429) * t + f32(-1.453152027)
    for ck in (1.421413741, -0.284496736, 0.254829592):
        poly = (poly * t + f32(ck)).astype(f32)
    ga = (x * (0.5 * (1 + np.copysign((1 - poly * t * ee).astype(f32), x))).astype(f32)).astype(f32)
    print("A&S 7.1.26:  max |gelu err| / max(1,|x|) = %.3e" % (np.abs(ga - gt) / np.maximum(1, np.abs(xd))).max())


if __name__ == "__main__":
    main()
