"""Tree-size statistics of the radon-919 bench workload (how predictable is the deepest tree of a lock-step round?).

  python tools/tree_stats.py [chains] [burn_in] [draws]   ->  gpurun_out/tree_stats.npz + a printed summary
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pymc4_b200._cabi import DeviceModel  # noqa: E402
from pymc4_b200._cstructs import PM4_EPS_POOLED, make_adapt, make_nuts_cfg  # noqa: E402
from pymc4_b200.ir import lower_model  # noqa: E402
from pymc4_b200 import benchmodels as M  # noqa: E402


def main():
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    idx, x, y, g = M.radon_synthetic(919, 85)
    ir, _, _ = lower_model(M.hierarchical_model(idx, x, y, g))
    dm = DeviceModel(ir, device=0)
    cfg = make_nuts_cfg(C, S, B, step_size=0.1, seed=7, adapt=make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=B))
    out = dm.nuts(cfg, np.zeros((C, ir.D), np.float32), dtype=np.float32)
    lf = out["leapfrogs"].cpu().numpy()  # [S, C]
    dv = out["diverging"].cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "tree_stats.npz"), leapfrogs=lf, diverging=dv)
    print("eps", float(out["step_size"][0]), "mean leapfrogs", lf.mean())
    b = np.floor(np.log2(np.maximum(lf, 1))).astype(int)
    print("bucket histogram", np.bincount(b.ravel(), minlength=11) / b.size)
    print("per-draw max: histogram of bucket of the max", np.bincount(b.max(axis=1), minlength=11))
    # P(bucket_t | bucket_{t-1})
    T = np.zeros((11, 11))
    np.add.at(T, (b[:-1].ravel(), b[1:].ravel()), 1)
    with np.printoptions(precision=3, suppress=True, linewidth=200):
        print("transition matrix rows=prev bucket")
        print(T / np.maximum(T.sum(1, keepdims=True), 1))


if __name__ == "__main__":
    main()
