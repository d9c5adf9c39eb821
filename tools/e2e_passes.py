"""Per-pass host-side breakdown of the end-to-end NUTS call of the bench (radon-919, 4096 chains, B = S = 1000):
which of the calls under `NUTS.__call__` the pass-to-pass spread of the e2e time sits in.

  python tools/e2e_passes.py [passes]
"""
import cProfile
import gc
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pymc4_b200 import _cabi  # noqa: E402
from pymc4_b200 import benchmodels as M  # noqa: E402
from pymc4_b200.mcmc.samplers import NUTS  # noqa: E402

WATCH = ("_pinned", "nuts", "synchronize", "_fetch_all", "trace_to_arviz", "lower_model", "_get_device_model",
         "trace_fn", "_run_chains", "_sample", "to_device", "empty", "collect", "pm4_nuts_run")


def main():
    passes = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    model = M.hierarchical_model(*M.radon_synthetic())
    C, B, S = 4096, 1000, 1000
    nuts = NUTS(model, step_size=0.1, num_adaptation_steps=B, device=0)
    pinned_log = []
    orig_pinned = _cabi.DeviceModel._pinned

    def pinned(self, shape, dtype):
        t0 = time.perf_counter()
        before = id(self.__dict__.get("_pinned_pool", {}).get((tuple(int(v) for v in shape), dtype)))
        r = orig_pinned(self, shape, dtype)
        after = id(self.__dict__["_pinned_pool"][(tuple(int(v) for v in shape), dtype)])
        pinned_log.append(("hit" if before == after else "MISS", 1e3 * (time.perf_counter() - t0)))
        return r

    _cabi.DeviceModel._pinned = pinned
    trace = None
    for k in range(passes):
        trace = None
        gc.collect()
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        trace = nuts(num_samples=S, num_chains=C, burn_in=B, seed=5 + k)
        torch.cuda.synchronize()
        pr.disable()
        dt = 1e3 * (time.perf_counter() - t0)
        st = pstats.Stats(pr)
        rows = []
        for (fn, line, name), (cc, nc, tt, ct, callers) in st.stats.items():
            if name in WATCH or any(w in name for w in ("pm4_", "synchronize")):
                if ct >= 2e-3:
                    rows.append((ct, tt, nc, "%s:%d %s" % (os.path.basename(fn), line, name)))
        rows.sort(reverse=True)
        print("pass %d: %.1f ms  pinned=%s" % (k, dt, pinned_log[-1] if pinned_log else None))
        for ct, tt, nc, what in rows[:14]:
            print("    cum %8.1f ms  self %8.1f ms  calls %4d  %s" % (1e3 * ct, 1e3 * tt, nc, what))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
