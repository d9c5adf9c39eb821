"""radon-919: reference-default shared step size against diagonal mass-matrix windows (R-hat, divergences, tree size)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymc4_b200 import benchmodels as M, diagnostics  # noqa: E402
from pymc4_b200._cabi import DeviceModel  # noqa: E402
from pymc4_b200._cstructs import PM4_EPS_POOLED, make_adapt, make_nuts_cfg  # noqa: E402
from pymc4_b200.ir import lower_model  # noqa: E402

ir, _, _ = lower_model(M.hierarchical_model(*M.radon_synthetic()))
dm = DeviceModel(ir, device=0)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
cols = [0, 1, 2, 87, 172, 173, 174]
from pymc4_b200._cstructs import PM4_EPS_PER_CHAIN  # noqa: E402

for K, mode, ta in ((0, PM4_EPS_POOLED, 0.75), (4, PM4_EPS_POOLED, 0.75), (4, PM4_EPS_POOLED, 0.9), (4, PM4_EPS_POOLED, 0.95),
                    (0, PM4_EPS_PER_CHAIN, 0.75), (4, PM4_EPS_PER_CHAIN, 0.75), (4, PM4_EPS_PER_CHAIN, 0.9)):
    print("mode", "pooled" if mode == PM4_EPS_POOLED else "per-chain", "target accept", ta, end=" | ")
    cfg = make_nuts_cfg(C, S, B, step_size=0.1, seed=12, adapt=make_adapt(mode=mode, num_adaptation_steps=B, mass_windows=K, target_accept_prob=ta))
    out = dm.nuts(cfg, np.zeros((C, ir.D), np.float32), dtype=np.float32)
    summ = diagnostics.summarize(out["states"], columns=cols)
    dv = out["diverging"].float()
    stuck = float((dv.mean(dim=0) > 0.9).float().mean())
    ms = out["mass_scale"].cpu().numpy() if "mass_scale" in out else np.ones(ir.D)
    print("K=%d eps %.4f leapfrogs/draw %.1f divergent %.4f stuck chains %.4f max split R-hat %.3f min ESS %.0f accept %.3f scale[mu_a,a0,b0,ls_a,ls_b,l_eps] %s"
          % (K, float(out["step_size"][0]), float(out["leapfrogs"].float().mean()), float(dv.mean()), stuck, summ["max_rhat"],
             summ["min_ess"], float(np.exp(out["log_accept"].cpu().numpy()).mean()), np.round(ms[[0, 2, 87, 172, 173, 174]], 2)))
