"""Fit the polynomial used by the GEMM epilogue's erf-GELU (csrc/common.cuh: gelu2) and report its error.

gelu(x) = relu(x) - |x/2| * erfc(|x|/sqrt2); log2(erfc(a/sqrt2)) is smooth on [0,6] and is fitted by a degree-6
polynomial (weighted so that the ABSOLUTE gelu error is minimised); beyond 6 the argument is clamped (erfc < 2e-9)."""
import math
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as Pn
from scipy.special import erf, erfc

A, DEG = 6.0, 6
a = np.linspace(0, A, 200001)
f = np.log2(erfc(a / np.sqrt(2)))
w = np.sqrt(0.5 * a * erfc(a / np.sqrt(2)) + 1e-6)
p = C.Chebyshev.fit(a, f, DEG, w=w, domain=[0, A]).convert(kind=Pn.Polynomial).coef.astype(np.float32)
print("coefficients P0..P6:", [repr(float(c)) for c in p])
x = np.linspace(-8, 8, 400001).astype(np.float32)
aa = np.minimum(np.abs(x), np.float32(A))
acc = np.full_like(aa, p[-1])
for c in p[-2::-1]:
    acc = (acc * aa + np.float32(c)).astype(np.float32)
g = (np.maximum(x, 0) - np.abs(np.float32(0.5) * x) * np.exp2(acc).astype(np.float32)).astype(np.float32)
ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / math.sqrt(2)))
print("max |gelu error| on [-8,8]:", float(np.abs(g - ref).max()))
