"""Fit used by csrc/common.cuh::gelu_erf_f: degree-6 minimax (Lawson-weighted least squares) polynomial of
q(a) = log2(Phi(a)) on a in [-5.5, 0]; gelu(x) = x * (x > 0 ? 1 - 2^q(-|x|) : 2^q(-|x|)).  Prints the coefficients
(lowest order first) and the error of the fp32 Horner evaluation against the exact-erf GELU."""
import numpy as np
from scipy.special import erf, log_ndtr

A, n, N = 5.5, 6, 6000
t = np.cos(np.pi * (np.arange(N) + 0.5) / N)
a = (t - 1) / 2 * A
q = log_ndtr(a) / np.log(2)
V = np.stack([(a / A) ** i for i in range(n + 1)], 1)
w = np.ones(N)
for _ in range(80):
    c, *_ = np.linalg.lstsq(V * w[:, None], q * w, rcond=None)
    e = np.abs(V @ c - q)
    w = w * (e / e.max() + 1e-3) ** 0.5
    w /= w.max()
c = c / (A ** np.arange(n + 1))
print("coefficients:", [float(x) for x in c])
x = np.linspace(-8, 8, 800001)
xf = np.float32(x)
af = np.maximum(-np.abs(xf), np.float32(-A))
p = np.full_like(af, np.float32(c[-1]))
for ci in c[-2::-1]:
    p = (p * af + np.float32(ci)).astype(np.float32)
phin = np.exp2(p).astype(np.float32)
gel = xf * np.where(xf > 0, np.float32(1) - phin, phin)
ref = 0.5 * x * (1 + erf(x / np.sqrt(2)))
err = np.abs(gel - ref)
print("max |err|", err.max(), " max rel err (|ref| > 1e-3)", (err / np.abs(ref))[np.abs(ref) > 1e-3].max())
