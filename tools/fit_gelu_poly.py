"""Coefficients of gelu_erf_fast (maest_b200/csrc/gemm.cuh): erfc(|x|/sqrt 2) = 2^q(|x|), q = degree-7 polynomial, q(0) = 0.

Weighted least squares on Chebyshev nodes of z = |x|/sqrt 2 in [0, 4]; the weight asks for <= 1e-6 absolute error of erf
near 0 and <= 2e-4 relative error of erfc in the tail.  Prints the float32 coefficients and the achieved GELU error."""
import numpy as np
from scipy.special import erf, erfc

Z, DEG = 4.0, 7
z = np.cos(np.linspace(0, np.pi, 8001)) * Z / 2 + Z / 2
w = np.maximum(erfc(z) / 1e-6, 1 / 2e-4)
a = z * np.sqrt(2)
V = np.vander(a, DEG + 1, increasing=True)[:, 1:]
c = np.linalg.lstsq(V * w[:, None], np.log2(erfc(z)) * w, rcond=None)[0].astype(np.float32)
print("q(|x|) = |x| * (c1 + |x| * (c2 + ...)):", [f"{v:.9e}" for v in c])
x = np.linspace(-Z * np.sqrt(2), Z * np.sqrt(2), 400001).astype(np.float32)
ax = np.abs(x)
acc = np.full_like(ax, c[-1])
for k in range(len(c) - 2, -1, -1):
    acc = acc * ax + c[k]
e = np.exp2((acc * ax).astype(np.float64))
hx = 0.5 * x.astype(np.float64)
g = hx + np.abs(hx) * (1 - e)
gt = hx * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
print("max |gelu err|", np.abs(g - gt).max(), " max rel err where |gelu| > 1e-4:", (np.abs(g - gt) / np.maximum(np.abs(gt), 1e-4)).max())
