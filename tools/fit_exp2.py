"""Degree-3 minimax polynomial for 2^r on [-0.5, 0.5] used by ex2_poly_pair (csrc/attention_cs.cu, the VTC_ACS_POLY experiment)
+ a check of the bit-trick reconstruction 2^x = 2^round(x) * p(x - round(x)) in fp32 arithmetic."""
import numpy as np
from scipy.optimize import minimize

r = np.cos(np.linspace(0, np.pi, 4001)) * 0.5
f = 2.0 ** r
A = np.stack([r ** k / f for k in range(4)], 1)          # relative error p(r) / 2^r - 1 is linear in the coefficients
err = lambda c: A @ c - 1
c = np.linalg.lstsq(A, np.ones_like(r), rcond=None)[0]
for _ in range(200):                                       # re-weighted least squares towards the minimax fit, then a polish
    e = err(c)
    w = (np.abs(e) / np.abs(e).max()) ** 2 + 1e-3
    c = np.linalg.lstsq(A * w[:, None], w, rcond=None)[0]
c = minimize(lambda c: np.abs(err(c)).max(), c, method="Nelder-Mead", options=dict(xatol=1e-12, fatol=1e-14, maxiter=20000)).x
p32 = c.astype(np.float32)
print("coefficients (c0..c3):", [repr(float(v)) for v in p32])
x = np.linspace(-120, 8, 2000001).astype(np.float32)
magic = np.float32(12582912.0)
t = (x + magic).astype(np.float32)
fr = (x - (t - magic)).astype(np.float32)
acc = np.full_like(fr, p32[-1])
for v in p32[-2::-1]:
    acc = (acc * fr + v).astype(np.float32)
e = (acc.view(np.uint32) + (t.view(np.uint32) << np.uint32(23))).view(np.float32)
ref = np.exp2(x.astype(np.float64))
print("max relative error on [-120, 8]:", float((np.abs(e - ref) / ref).max()))
