"""Fit of the bf16-mode GELU used by csrc/mixffn_tc.cuh:
   gelu(x) ~= x * sigmoid(x (a + b x^2 + c x^4)),  minimax over [-9, 9] against the erf form."""
import math

import numpy as np
from scipy.optimize import minimize

x = np.linspace(-9, 9, 600001)
exact = 0.5 * x * (1 + np.vectorize(math.erf)(x / math.sqrt(2)))


def approx(p):
    return x / (1 + np.exp(-x * (p[0] + p[1] * x ** 2 + p[2] * x ** 4)))


r = minimize(lambda p: np.max(np.abs(approx(p) - exact)), [1.5950158, 0.0740113, -0.000703],
             method='Nelder-Mead', options=dict(xatol=1e-10, fatol=1e-13, maxiter=8000))
print('a, b, c =', r.x, ' max abs error', r.fun)
# the kernel uses the equivalent tanh form 0.5 x (1 + tanh(x (a/2 + b/2 x^2 + c/2 x^4)))
print('tanh-form coefficients:', [v / 2 for v in r.x])
print('ex2 form coefficients (x -log2 e):', [-v * 1.4426950408889634 for v in r.x])
