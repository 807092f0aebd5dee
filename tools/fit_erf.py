"""Weighted-minimax fit behind gelu_erf() in vss_cffm_b200/csrc/common.cuh:
erf(z) = 1 - 2^(-z P5(z)) on [0, 4] (erf(4) = 1 - 1.5e-8), weight erfc(z) so that the ABSOLUTE error of erf
is minimised; evaluated in float32 with Horner.  Prints the coefficients and the max abs error."""
import numpy as np
from scipy.special import erf, erfc

D, zmax = 5, 4.0
k = np.arange(4000)
z = 0.5 * zmax * (1 - np.cos(np.pi * (k + 0.5) / 4000))
z = z[z > 1e-6]
q = -np.log2(erfc(z))
w = erfc(z) * np.log(2)
A = np.stack([z ** (i + 1) for i in range(D + 1)], 1)
wt = w.copy()
for _ in range(60):
    c, *_ = np.linalg.lstsq(A * wt[:, None], q * wt, rcond=None)
    err = np.abs((A @ c - q) * w)
    wt = wt * (1 + 2 * err / err.max())
zz = np.linspace(0, 6, 200001).astype(np.float32)
zc = np.minimum(zz, np.float32(zmax))
c32 = c.astype(np.float32)
p = np.full_like(zc, c32[-1])
for ci in c32[-2::-1]:
    p = p * zc + ci
e = 1 - np.exp2(-(p * zc))
print("coefficients (z^1 .. z^6):", c)
print("max abs error (fp32):", np.abs(e.astype(np.float64) - erf(zz.astype(np.float64))).max())
