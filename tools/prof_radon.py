"""Profiling helper: one NUTS run of the synthetic radon-919 model (the bench workload) sized for ncu.

  python tools/prof_radon.py [chains] [burn_in] [draws] [pooled|per-chain]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pymc4_b200._cabi import DeviceModel  # noqa: E402
from pymc4_b200._cstructs import PM4_EPS_PER_CHAIN, PM4_EPS_POOLED, make_adapt, make_nuts_cfg  # noqa: E402
from pymc4_b200.ir import lower_model  # noqa: E402
from pymc4_b200 import benchmodels as M  # noqa: E402


def main():
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    mode = PM4_EPS_POOLED if (len(sys.argv) > 4 and sys.argv[4] == "pooled") else PM4_EPS_PER_CHAIN
    eps0 = float(sys.argv[5]) if len(sys.argv) > 5 else 0.03
    idx, x, y, g = M.radon_synthetic(919, 85)
    ir, _, _ = lower_model(M.hierarchical_model(idx, x, y, g))
    dm = DeviceModel(ir, device=0)
    cfg = make_nuts_cfg(C, S, B, step_size=eps0, seed=7, adapt=make_adapt(mode=mode, num_adaptation_steps=B))
    import torch

    out = dm.nuts(cfg, np.zeros((C, ir.D), np.float32), dtype=np.float32)
    torch.cuda.synchronize()
    print("leapfrogs", int(out["total_leapfrogs"].sum().item()), "eps", float(out["step_size"][0]))


if __name__ == "__main__":
    main()
