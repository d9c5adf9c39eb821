"""Smoke-sized invocations of every kernel family for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pymc4_b200._cabi import DeviceModel  # noqa: E402
from pymc4_b200._cstructs import (PM4_EPS_PER_CHAIN, PM4_EPS_POOLED, make_adapt, make_hmc_cfg,  # noqa: E402
                                  make_nuts_cfg)
from pymc4_b200.ir import lower_model  # noqa: E402
from tests import models as M  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    # warp-per-chain NUTS (owner-computes layout): lock-step rounds (cooperative launch), mass windows, persistent blocks
    ir, _, _ = lower_model(M.radon_model(real=False))
    dm = DeviceModel(ir, device=0)
    th = rng.normal(0, 0.2, (16, ir.D)).astype(np.float32)
    dm.logp_grad(th, dtype=np.float32)
    for mode, K in ((PM4_EPS_POOLED, 0), (PM4_EPS_POOLED, 2), (PM4_EPS_PER_CHAIN, 0)):
        cfg = make_nuts_cfg(40, 6, 6, step_size=0.03, seed=1, adapt=make_adapt(mode=mode, num_adaptation_steps=6, mass_windows=K))
        dm.nuts(cfg, np.zeros((40, ir.D), np.float32), dtype=np.float32)
    dm.hmc(make_hmc_cfg(16, 4, 4, step_size=0.01, num_leapfrog_steps=3, seed=2, adapt=make_adapt(num_adaptation_steps=4)),
           np.zeros((16, ir.D), np.float32), dtype=np.float32)
    dm.log_likelihood(th, dtype=np.float32)
    dm.posterior_predictive(th, seed=3, dtype=np.float32)
    dm.prior_predictive(8, seed=4, dtype=np.float32)
    # warp path without a gather term (eight schools) and the legacy xs / part path (logistic: a general term)
    for mdl in (M.schools_pm4(), M.logistic_model(*M.logistic_data())):
        ir2, _, _ = lower_model(mdl)
        d2 = DeviceModel(ir2, device=0)
        d2.nuts(make_nuts_cfg(8, 5, 5, step_size=0.1, seed=5, adapt=make_adapt(num_adaptation_steps=5)),
                np.zeros((8, ir2.D), np.float32), dtype=np.float32)
    # dense GLM term (tcgen05) and marginal GP term (FP64 MMA), lock-step drivers
    irg, _, _ = lower_model(M.logistic_glm_model(*M.logistic_glm_data(n=512, p=12)))
    dg = DeviceModel(irg, device=0)
    dg.hmc(make_hmc_cfg(8, 2, 2, step_size=0.01, num_leapfrog_steps=2, seed=6, adapt=make_adapt(num_adaptation_steps=2)),
           np.zeros((8, irg.D), np.float32), dtype=np.float32)
    irp, _, _ = lower_model(M.gp_model(*M.gp_data(n=96)))
    dp = DeviceModel(irp, device=0)
    dp.hmc(make_hmc_cfg(4, 2, 2, step_size=0.005, num_leapfrog_steps=2, seed=7, adapt=make_adapt(num_adaptation_steps=2)),
           np.zeros((4, irp.D), np.float32), dtype=np.float32)
    torch.cuda.synchronize()
    print("sanitize smoke ok")


if __name__ == "__main__":
    main()
