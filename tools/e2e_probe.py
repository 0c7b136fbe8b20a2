#!/usr/bin/env python
"""Time the host-buffer entry point wt_step_host alone (ms per call, 1,048,576 x 10 config5 plants)."""
import ctypes as C, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ics_wt_physicsengine_b200 import _lib, ensembles  # noqa: E402
from ics_wt_physicsengine_b200.params import derive_params  # noqa: E402
P, n = int(os.environ.get("PLANTS", 1048576)), 10
e = ensembles.config5(P, n)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
par, bnd = pin(derive_params(e.cfg, n).T), pin(e.bnd.T)
y = pin(np.stack([e.pH0.T, e.Cl0.T, e.T0.T]))
t, flow = torch.zeros(P, dtype=torch.float64).pin_memory(), torch.zeros(P, dtype=torch.float64).pin_memory()
st = torch.zeros(P, dtype=torch.int32).pin_memory()
p = lambda x: C.c_void_p(x.data_ptr())
L = _lib.lib()
for i in range(4):
    _lib.check(L.wt_step_host(P, n, 1.0, p(par), p(bnd), P, p(t), p(y), p(flow), p(st), 64, 1 if i else 0), "wt_step_host")
t0 = time.perf_counter()
K = 10
for _ in range(K):
    _lib.check(L.wt_step_host(P, n, 1.0, p(par), p(bnd), P, p(t), p(y), p(flow), p(st), 64, 1), "wt_step_host")
el = (time.perf_counter() - t0) / K
print(f"slabs={os.environ.get('WT_B200_HOST_SLABS', 'default')}: {el * 1e3:.2f} ms per call, {P * n / el:.3e} plant-zone-steps/s", flush=True)
