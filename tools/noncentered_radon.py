"""Check (c) on a configuration that converges: the radon-919 workload with the group effects written as offsets
(benchmodels.hierarchical_model_noncentered), reference-default sampler settings (scalar step_size 0.1, pooled dual
averaging over burn_in, max_tree_depth 10).  Prints one JSON line: throughput, ESS/s, split R-hat, divergences, and the
agreement of the posterior means of the shared scalars with the centred model's.

  python tools/noncentered_radon.py [chains] [burn_in] [draws]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from pymc4_b200 import benchmodels as M, diagnostics  # noqa: E402
from pymc4_b200._cabi import DeviceModel  # noqa: E402
from pymc4_b200._cstructs import PM4_EPS_POOLED, make_adapt, make_nuts_cfg  # noqa: E402
from pymc4_b200.ir import lower_model  # noqa: E402


def run(model_fn, C, B, S, seed):
    idx, x, y, g = M.radon_synthetic(919, 85)
    ir, _, _ = lower_model(model_fn(idx, x, y, g))
    dm = DeviceModel(ir, device=0)
    cfg = make_nuts_cfg(C, S, B, step_size=0.1, seed=seed, adapt=make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=B))
    dm.nuts(cfg, np.zeros((C, ir.D), np.float32), dtype=np.float32)  # warm-up (module load, allocator)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = dm.nuts(cfg, np.zeros((C, ir.D), np.float32), dtype=np.float32)
    e1.record()
    e1.synchronize()
    secs = e0.elapsed_time(e1) * 1e-3
    cols = [0, 1, 2, 87, 172, 173, 174]
    summ = diagnostics.summarize(out["states"], columns=cols, gather=False)
    st = out["states"]
    means = {"mu_alpha": float(st[:, :, 0].mean()), "mu_beta": float(st[:, :, 1].mean()),
             "log_sigma_alpha": float(st[:, :, 172].mean()), "log_sigma_beta": float(st[:, :, 173].mean()),
             "log_eps": float(st[:, :, 174].mean())}
    return dict(grad_evals_per_sec=float(out["total_leapfrogs"].sum().item()) / secs, ms=1e3 * secs,
                ess_per_sec=summ["min_ess"] / secs, min_bulk_ess=summ["min_ess"], max_split_rhat=summ["max_rhat"],
                divergent_frac=float(out["diverging"].float().mean().item()),
                mean_leapfrogs_per_draw=float(out["leapfrogs"].float().mean().item()),
                step_size=float(out["step_size"][0].item()), means=means)


def main():
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    res = {"chains": C, "burn_in": B, "draws": S,
           "noncentered": run(M.hierarchical_model_noncentered, C, B, S, 11),
           "centered": run(M.hierarchical_model, C, B, S, 11)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
