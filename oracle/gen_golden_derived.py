#!/usr/bin/env python
"""Golden vectors of the DERIVED state (_update_derived_state, reactor.py:511-524), produced by RUNNING THE
UNMODIFIED REFERENCE: after every step() of 64 config-3 plants (n = 20; temperature profiles over 0-100 C, incl. the
<= 8 C density branch) the reference's state.H_concentration, state.density and state.chlorine_decay_rate are stored
next to the primary state.  Build container only.   python oracle/gen_golden_derived.py -> tests/golden/derived_config3.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as gg  # noqa: E402  (imports the reference, defines make_plant / ref_params)
from ics_wt_physicsengine_b200 import ensembles as ens  # noqa: E402


def main():
    P, n, steps = 64, 20, 4
    e = ens.config3(P, n, seed=20260009)
    Y = np.full((steps, P, 3 * n), np.nan)
    D = np.full((steps, P, 3 * n), np.nan)
    raised = np.full(P, -1, dtype=np.int32)
    for p in range(P):
        r, b = gg.make_plant(e, p)
        for s in range(steps):
            try:
                st = r.step(1.0, b)
            except ValueError:
                raised[p] = s
                break
            Y[s, p] = np.concatenate([st.pH, st.chlorine, st.temperature])
            D[s, p] = np.concatenate([st.H_concentration, st.density, st.chlorine_decay_rate])
    np.savez_compressed(os.path.join(gg.GOLD, "derived_config3.npz"), n_zones=n, dt=1.0, nsteps=steps, cfg=e.cfg, bnd=e.bnd,
                        pH0=e.pH0, Cl0=e.Cl0, T0=e.T0, Y=Y, D=D, raised=raised, **gg.STAMP)
    cold = (Y[:, :, 2 * n:] <= 8.0).any(axis=(0, 2)).sum()
    print("wrote", Y.shape, "raised:", int((raised >= 0).sum()), "plants with a zone <= 8 C:", int(cold))


if __name__ == "__main__":
    main()
