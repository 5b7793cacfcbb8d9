"""Design check for gvl::gelu_erf (csrc/ptx.cuh): exhaustive comparison over ALL finite bf16 inputs of
    gelu(x) = max(x, 0) - |x| * exp2(P7(min(|x|, A)))          (fp32 arithmetic, result rounded to bf16)
against double-precision x * 0.5 * erfc(-x / sqrt 2), torch's own fp32 GELU and the round-1 Abramowitz-Stegun form.
    python tools/check_gelu.py            # prints the fit, the mismatch counts and the coefficients used in ptx.cuh
CPU only (numpy / scipy / torch)."""
import math

import numpy as np
import torch
from scipy.special import erfc

A, DEG = 6.5, 7


def fit():
    a = np.linspace(0, A, 40001)
    L = np.log2(0.5 * erfc(a / math.sqrt(2)))
    p = np.polynomial.chebyshev.Chebyshev.fit(a, L, DEG, domain=[0, A]).convert(kind=np.polynomial.Polynomial)
    return p.coef.astype(np.float32), float(np.abs(p(a) - L).max())


def gelu_new(xf, coef):
    xf = xf.astype(np.float32)
    aa = np.minimum(np.abs(xf), np.float32(A))
    acc = np.full_like(aa, coef[-1])
    for k in range(len(coef) - 2, -1, -1):
        acc = (acc * aa + coef[k]).astype(np.float32)
    q = np.exp2(acc).astype(np.float32)
    return (np.maximum(xf, np.float32(0)) - np.abs(xf) * q).astype(np.float32)


def main():
    bits = np.arange(65536, dtype=np.uint32)
    x = torch.from_numpy((bits << 16).view(np.float32).copy())
    x = x[torch.isfinite(x)]
    xd = x.double().numpy()
    exact = xd * 0.5 * erfc(-xd / math.sqrt(2.0))
    bf = lambda v: torch.from_numpy(np.asarray(v, dtype=np.float32)).to(torch.bfloat16)
    ref = bf(exact)
    tg = torch.nn.functional.gelu(x.float()).to(torch.bfloat16)
    coef, lerr = fit()
    with np.errstate(over="ignore", invalid="ignore"):
        newf = gelu_new(xd, coef)
    new = bf(newf)
    body = torch.from_numpy(np.abs(xd) <= 1e4)
    d = (new.view(torch.int16).int() - ref.view(torch.int16).int()).abs()
    vis = body & torch.from_numpy(np.abs(exact) > 1e-6)
    print("fit: degree %d on [0, %.1f], max |log2 err| %.2e (relative error of q %.2e)" % (DEG, A, lerr, lerr * math.log(2)))
    print("finite bf16 inputs: %d" % len(x))
    print("bf16 outputs that differ from the correctly rounded result: new %d (|x| <= 1e4: %d; more than 1 ulp among |gelu| > 1e-6: %d); "
          "torch fp32 gelu %d" % (int((new != ref).sum()), int((new != ref)[body].sum()), int((d > 1)[vis].sum()), int((tg != ref).sum())))
    m8 = np.abs(xd) <= 8
    print("max abs error (before the bf16 rounding) for |x| <= 8: %.3g ; for 8 < |x| <= 1e4: %.3g" % (
        np.abs(newf.astype(np.float64) - exact)[m8].max(), np.abs(newf.astype(np.float64) - exact)[(~m8) & body.numpy()].max()))
    print("coefficients (c0 .. c%d):" % DEG, ", ".join("%.9ef" % v for v in coef))


if __name__ == "__main__":
    main()
