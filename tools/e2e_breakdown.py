"""Where the end-to-end time of one NUTS call goes (host side): python tools/e2e_breakdown.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pymc4_b200 import benchmodels as M  # noqa: E402
from pymc4_b200.ir import lower_model  # noqa: E402
from pymc4_b200.mcmc.samplers import NUTS  # noqa: E402

model = M.hierarchical_model(*M.radon_synthetic())
C, B, S = 4096, 1000, 1000
nuts = NUTS(model, step_size=0.1, num_adaptation_steps=B, device=0)
for k in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ir, _, _ = lower_model(model)
    t1 = time.perf_counter()
    tr = None
    tr = nuts(num_samples=S, num_chains=C, burn_in=B, seed=5 + k)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("pass %d: lower_model alone %.1f ms; whole call %.1f ms" % (k, 1e3 * (t1 - t0), 1e3 * (t2 - t1)))

import cProfile  # noqa: E402
import pstats  # noqa: E402

pr = cProfile.Profile()
pr.enable()
tr = None
tr = nuts(num_samples=S, num_chains=C, burn_in=B, seed=99)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
