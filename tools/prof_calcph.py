#!/usr/bin/env python
"""Driver for ncu captures of the calculate_pH kernel: the 262,144 (+29) solves of BASELINE configs[3]."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ics_wt_physicsengine_b200 import calculate_pH_batch, ensembles  # noqa: E402

t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
alk, ct, temp, guess = (t(x) for x in ensembles.config4(262144))
ms = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ph, it, st = calculate_pH_batch(alk, ct, temp, guess)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
print("ms per launch:", ["%.3f" % m for m in ms])
print("iterations:", int(it.sum()), "status counts:", torch.bincount(st.to(torch.int64)).tolist())
