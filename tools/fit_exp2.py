"""Degree-5 polynomial for 2^f on [-0.5, 0.5] used by exp2_poly2 (csrc/common.cuh) + the bit-trick reconstruction check."""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as Pn

f = np.linspace(-0.5, 0.5, 200001)
p32 = C.Chebyshev.fit(f, np.exp2(f), 5, domain=[-0.5, 0.5]).convert(kind=Pn.Polynomial).coef.astype(np.float32)
print("coefficients:", [repr(float(c)) for c in p32])
x = np.linspace(-120, 8, 2000001).astype(np.float32)
magic = np.float32(12582912.0)
t = (x + magic).astype(np.float32)
fr = (x - (t - magic)).astype(np.float32)
acc = np.full_like(fr, p32[-1])
for c in p32[-2::-1]:
    acc = (acc * fr + c).astype(np.float32)
e = (acc.view(np.uint32) + (t.view(np.uint32) << np.uint32(23))).view(np.float32)
ref = np.exp2(x.astype(np.float64))
print("max relative error on [-120, 8]:", float((np.abs(e - ref) / ref).max()))
