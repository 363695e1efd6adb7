"""Trains one real element (N from argv, default 2048) a few times: target for `ncu` launch lists of the train path."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_liouville_equation_b200 import _lib as L
from gaussian_process_liouville_equation_b200 import synthetic as syn

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ctx = L.Context(0)
X, y = syn.training_set(70, 0, N, (0.0, syn.P0))
yv = np.ascontiguousarray(y).view(np.float64)
th = syn.theta_real()
for _ in range(3):
    h, sc = C.c_void_p(), L.RealScalars()
    ctx.check(ctx.lib.gple_train_real(ctx.h, L.addr(X), L.addr(yv), N, L.addr(th), 3, C.byref(h), C.byref(sc)))
    ctx.lib.gple_model_destroy(ctx.h, h)
print("ok", sc.error, sc.population)
