"""Fit of the logistic-form GELU used by the ALU-bound epilogues (csrc/common.cuh::gelu_fast):
   Phi(x) = 0.5 (1 + erf(x / sqrt 2)) ~= 1 / (1 + 2^(x (q0 + q1 x^2 + q2 x^4))) on |x| <= 4.75, iteratively reweighted
   least squares towards minimax.  Prints the coefficients and the fp32-evaluated error of gelu(x) = x Phi(x)."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf

x = np.linspace(0, 4.75, 40001)
Phi = 0.5 * (1 + erf(x / np.sqrt(2)))


def arg(q, x):
    x2 = x * x
    return x * (q[0] + x2 * (q[1] + x2 * q[2]))


def model(q, x):
    return 1.0 / (1.0 + np.exp2(arg(q, x)))


q = np.array([-2.3017, -0.1062, 0.0])
w = np.ones_like(x)
for _ in range(300):
    q = least_squares(lambda q: (model(q, x) - Phi) * w, q, xtol=1e-15, ftol=1e-15, gtol=1e-15).x
    e = np.abs(model(q, x) - Phi)
    w = w * (1 + 2 * e / e.max())
    w /= w.mean()
print("q =", list(q), "max |Phi err| =", np.abs(model(q, x) - Phi).max())
f32 = np.float32
xs = np.linspace(-10, 10, 2000001).astype(f32)
xc = np.clip(xs, -7, 7).astype(f32)
x2 = (xc * xc).astype(f32)
qq = q.astype(f32)
p = (qq[0] + x2 * (qq[1] + x2 * qq[2]).astype(f32)).astype(f32)
out = (xs * (f32(1) / (f32(1) + np.exp2((xc * p).astype(np.float64)).astype(f32)))).astype(f32)
ref = 0.5 * xs.astype(np.float64) * (1 + erf(xs.astype(np.float64) / np.sqrt(2)))
print("gelu max abs err (fp32 eval, all x) =", np.abs(out - ref).max())
