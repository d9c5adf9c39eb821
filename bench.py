#!/usr/bin/env python
"""bench.py -- leapfrog gradient-evaluations/sec of the NUTS hot path on the radon-shaped model.

Workload (BASELINE.json configs[1]): synthetic radon hierarchical model, 919 observations /
85 counties (D = 175), 4096 chains of NUTS per GPU.  One "step" = one complete sampling call of
the hot path for that batch of chains: bootstrap + `burn_in` adaptive transitions + `draws`
kept transitions from the common test point.  Default step-size policy = pm.sample's own default, a
scalar `step_size` (one epsilon adapted from the mean acceptance of all chains of the rank);
`--eps-mode per-chain` = `step_size` with a leading chain axis (independent adaptation).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the CPU implementation on the host cores)

Rank 0 prints ONE JSON line.  value = leapfrog gradient evaluations of all chains of all ranks
divided by the device time (CUDA events, max over ranks).  e2e = the same metric through the
public sampler API with host buffers (H2D of inputs and D2H of the trace inside the timing).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_OBS, N_GROUPS, CHAINS_PER_GPU = 919, 85, 4096
METRIC = "leapfrog_grad_evals_per_sec"
UNIT = "grad-evals/s"
# algorithmic cost of one unit (one leapfrog gradient evaluation of one chain), SURVEY section 8(d):
# F = 8 N + 14 G + 15 D flops; shared-data bytes = 12 N (idx, x, y read from on-chip memory)
D_MODEL = 2 * N_GROUPS + 5
FLOPS_PER_UNIT = 8 * N_OBS + 14 * N_GROUPS + 15 * D_MODEL
ONCHIP_BYTES_PER_UNIT = 12 * N_OBS


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU, help="chains per GPU")
    ap.add_argument("--burn-in", type=int, default=1000)  # S = 1000, B = 1000
    ap.add_argument("--draws", type=int, default=1000)
    ap.add_argument("--step-size", type=float, default=0.1)
    ap.add_argument("--seed", type=int, default=20261019)
    ap.add_argument("--cpu-chains", type=int, default=0, help="chains of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--warps-per-block", type=int, default=0)
    ap.add_argument("--max-tree-depth", type=int, default=10)
    ap.add_argument("--eps-mode", default="pooled", choices=["per-chain", "pooled"],
                    help="per-chain: step_size with a leading chain axis (independent dual averaging, no coupling); "
                         "pooled: scalar step_size, the reference default (one epsilon adapted from the mean "
                         "acceptance of all chains of the rank: lock-step launches during burn-in)")
    ap.add_argument("--workload", default="radon919", choices=["radon919", "radon100k", "logit", "gp"],
                    help="radon919 = BASELINE configs[1] (the headline); radon100k = configs[4] shape "
                         "(100k observations / 5k groups, D = 10 005; pass --chains/--burn-in/--draws to bound the run); "
                         "logit = configs[2]: hierarchical logistic regression, N = 1M x P = 100, 1024 chains of HMC "
                         "(3 leapfrogs), X beta on the tcgen05 tensor cores in 3xTF32; "
                         "gp = configs[3]: marginal GP regression, n = 4096 points, 256 chains of HMC (3 leapfrogs), "
                         "float64 batched Cholesky")
    ap.add_argument("--gp-n", type=int, default=4096)
    ap.add_argument("--logit-n", type=int, default=1_000_000)
    ap.add_argument("--leapfrogs", type=int, default=3, help="HMC leapfrog steps (logit workload; reference default 3)")
    ap.add_argument("--mass-windows", type=int, default=0,
                    help="diagonal mass-matrix windows during burn-in (backend extension; 0 = the reference behaviour)")
    ap.add_argument("--couple-step-size", action="store_true",
                    help="N > 1, pooled step size: ONE epsilon for the chains of ALL ranks (the box runs one pm.sample call: "
                         "a 16-byte exchange over NVLink peer memory inside the sampler per burn-in transition).  Default: "
                         "every rank is its own pm.sample call on its shard of the chains -- the reference has one process "
                         "per device and pools the step size over that process's chains -- and no data-path collective")
    ap.add_argument("--no-secondary", action="store_true",
                    help="radon919 only: do not append the bounded runs of the other BASELINE configurations "
                         "(logit, gp, radon100k) as the line's \"secondary\" object")
    ap.add_argument("--shard-rank", type=int, default=-1,
                    help="(secondary runs) this single-GPU process is shard R of a box-wide run: chains get the "
                         "global offsets R * chains")
    return ap.parse_args()


def build_model():
    from pymc4_b200.ir import lower_model
    from pymc4_b200 import benchmodels as M

    if WORKLOAD == "gp":
        X, y = M.gp_data(n=N_OBS, d=1, seed=20261019)
        model = M.gp_model(X, y)
        ir, _, _ = lower_model(model)
        assert ir.D == D_MODEL and len(ir.gp_terms) == 1
        return model, ir
    if WORKLOAD == "logit":
        X, y = M.logistic_glm_data(n=N_OBS, p=LOGIT_P, seed=20261019)
        model = M.logistic_glm_model(X, y)
        ir, _, _ = lower_model(model)
        assert ir.D == D_MODEL and len(ir.glm_terms) == 1
        return model, ir
    if N_OBS == 919:
        idx, x, y, g = M.radon_synthetic(N_OBS, N_GROUPS)
    else:
        idx, x, y, g = M.radon_large(N_OBS, N_GROUPS)
    model = M.hierarchical_model(idx, x, y, g)
    ir, _, _ = lower_model(model)
    assert ir.D == D_MODEL
    return model, ir


WORKLOAD, LOGIT_P = "radon919", 100


def select_workload(name, logit_n=1_000_000, gp_n=4096):
    """Switch the module-level shape constants (the algorithmic cost per unit follows the shape)."""
    global N_OBS, N_GROUPS, D_MODEL, FLOPS_PER_UNIT, ONCHIP_BYTES_PER_UNIT, WORKLOAD
    WORKLOAD = name
    if name == "gp":
        # F ~ n^3 flops per unit (Cholesky n^3/3 + inverse factor n^3/3 + contraction with dK n^3/3)
        N_OBS, N_GROUPS = int(gp_n), 0
        D_MODEL = 3
        FLOPS_PER_UNIT = float(N_OBS) ** 3
        ONCHIP_BYTES_PER_UNIT = 0
        return
    if name == "logit":
        # F = 4 N P flops per unit (X beta and X^T r), X read once per lock-step leapfrog for all chains
        N_OBS, N_GROUPS = int(logit_n), 0
        D_MODEL = LOGIT_P + 2
        FLOPS_PER_UNIT = 4 * N_OBS * LOGIT_P
        ONCHIP_BYTES_PER_UNIT = 0
        return
    if name == "radon100k":
        N_OBS, N_GROUPS = 100_000, 5_000
    D_MODEL = 2 * N_GROUPS + 5
    FLOPS_PER_UNIT = 8 * N_OBS + 14 * N_GROUPS + 15 * D_MODEL
    ONCHIP_BYTES_PER_UNIT = 12 * N_OBS


def workload_config(args, n_gpus):
    if WORKLOAD == "gp":
        return {
            "workload": "marginal GP regression, synthetic n=%d points (1-D), ExpQuad + noise covariance, MvNormal "
                        "likelihood (D=3 hyper-parameters), %d chains HMC (%d leapfrogs) per GPU, float64 batched blocked "
                        "Cholesky" % (N_OBS, args.chains, args.leapfrogs),
            "chains_per_gpu": args.chains, "global_chains": args.chains * n_gpus, "burn_in": args.burn_in,
            "draws": args.draws, "num_leapfrog_steps": args.leapfrogs,
            "step": "one full sampling call: bootstrap + burn_in adaptive + draws kept HMC transitions, lock-step over chains",
            "step_size": "%s dual averaging from %.3g" % (args.eps_mode, args.step_size),
            "parallelism": "chains sharded, %d per GPU; every GPU holds X and y" % args.chains,
            "l2": "L2 flushed (256 MiB write) between timed steps; the per-chain factors (%.0f GB) exceed L2"
                  % (args.chains * 2 * N_OBS * N_OBS * 8 / 1e9),
        }
    if WORKLOAD == "logit":
        return {
            "workload": "hierarchical logistic regression, synthetic N=%d x P=%d Bernoulli likelihood (D=%d), %d chains "
                        "HMC (%d leapfrogs) per GPU, X.beta as 3xTF32 tcgen05 tiles" % (N_OBS, LOGIT_P, D_MODEL, args.chains,
                                                                                          args.leapfrogs),
            "chains_per_gpu": args.chains, "global_chains": args.chains * n_gpus, "burn_in": args.burn_in,
            "draws": args.draws, "num_leapfrog_steps": args.leapfrogs,
            "step": "one full sampling call: bootstrap + burn_in adaptive + draws kept HMC transitions, lock-step over chains",
            "step_size": "%s dual averaging from %.3g" % (args.eps_mode, args.step_size),
            "parallelism": "chains sharded, %d per GPU; every GPU holds the full X" % args.chains,
            "l2": "L2 flushed (256 MiB write) between timed steps; X (400 MB) exceeds L2",
        }
    return {
        "workload": "radon_hierarchical varying-intercept/slope, synthetic %d obs / %d %s (D=%d), "
                    "%d chains NUTS per GPU" % (N_OBS, N_GROUPS, "counties" if N_OBS == 919 else "groups", D_MODEL,
                                                args.chains),
        "chains_per_gpu": args.chains,
        "global_chains": args.chains * n_gpus,
        "burn_in": args.burn_in,
        "draws": args.draws,
        "step": "one full sampling call: bootstrap + burn_in adaptive + draws kept NUTS transitions",
        "step_size": ("per-chain dual averaging from %.3g (leading chain axis; no cross-chain coupling)" % args.step_size)
                     if args.eps_mode == "per-chain" else
                     (("pooled dual averaging from %.3g (scalar step_size, reference default; ONE epsilon for the chains "
                       "of all ranks: per-transition exchange over NVLink peer memory inside the sampler kernel)"
                       if (n_gpus > 1 and args.couple_step_size) else
                       "pooled dual averaging from %.3g (scalar step_size, reference default: one epsilon per pm.sample call "
                       "= per rank, adapted from the mean acceptance of that rank's chains)") % args.step_size),
        "max_tree_depth": args.max_tree_depth,
        "parallelism": ("chains sharded, %d per GPU; " % args.chains)
                       + ("no data-path collective (NCCL all-gather of the monitored columns for R-hat / ESS only)"
                          if (n_gpus == 1 or args.eps_mode == "per-chain" or not args.couple_step_size) else
                          "one 16-byte all-reduce per burn-in transition (pooled step size), none afterwards"),
        "l2": "L2 flushed (256 MiB write) between timed steps",
    }


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle (TFP-semantics restatement, oracle/pm4_oracle.c) on the host cores
# ------------------------------------------------------------------------------------------
def cpu_sample_rate(ir, args, chains, seed):
    from oracle import pm4_oracle
    from pymc4_b200._cstructs import PM4_EPS_PER_CHAIN, PM4_EPS_POOLED, make_adapt, make_nuts_cfg

    orc = pm4_oracle.Oracle(ir)
    mode = PM4_EPS_PER_CHAIN if args.eps_mode == "per-chain" else PM4_EPS_POOLED
    adapt = make_adapt(mode=mode, num_adaptation_steps=args.burn_in)
    if WORKLOAD == "gp":
        # one scalar gradient evaluation at n = 4096 is ~7e10 flops: the sample is one evaluation per chain
        rng = np.random.default_rng(seed)
        th = rng.normal(0, 0.1, (chains, ir.D))
        th[:, 2] -= 2.0
        t0 = time.perf_counter()
        orc.logp_grad(th, dtype=np.float64)
        dt = time.perf_counter() - t0
        return float(chains), dt
    if WORKLOAD == "logit":
        from pymc4_b200._cstructs import make_hmc_cfg

        # a scalar gradient evaluation over N = 1M rows takes ~0.2 s: bound the sample to 2 + 2 transitions
        # (the cost per gradient evaluation does not depend on the transition count)
        b, s_ = min(args.burn_in, 2), min(args.draws, 2)
        adapt = make_adapt(mode=mode, num_adaptation_steps=b)
        cfg = make_hmc_cfg(chains, s_, b, step_size=args.step_size, num_leapfrog_steps=args.leapfrogs,
                           seed=seed, adapt=adapt)
        t0 = time.perf_counter()
        orc.hmc(cfg, np.zeros(ir.D), dtype=np.float32)
        dt = time.perf_counter() - t0
        return float(chains * (b + s_) * args.leapfrogs), dt
    draws, burn = args.draws, args.burn_in
    if WORKLOAD == "radon100k":
        # a scalar leapfrog over 100k observations costs ~0.4 ms and early trees are hundreds of leapfrogs deep:
        # bound the sample to 10 + 2 transitions (the cost per gradient evaluation does not depend on the count)
        burn, draws = min(burn, 10), min(draws, 2)
        adapt = make_adapt(mode=mode, num_adaptation_steps=burn)
    cfg = make_nuts_cfg(chains, draws, burn, step_size=args.step_size, seed=seed, adapt=adapt)
    t0 = time.perf_counter()
    out = orc.nuts(cfg, np.zeros(ir.D), dtype=np.float32)
    dt = time.perf_counter() - t0
    return float(out["total_leapfrogs"].sum()), dt


def cpu_baseline(ir, args):
    from oracle import pm4_oracle

    cores = os.cpu_count() or 1
    pm4_oracle.build()
    cores = pm4_oracle.set_threads(cores)  # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every host thread
    chains = args.cpu_chains or cores
    # calibrate so the sample costs roughly 10-30 s of CPU work
    evals, dt = cpu_sample_rate(ir, args, chains, args.seed)
    rate = evals / dt
    if dt < 8.0 and not args.cpu_chains and WORKLOAD not in ("logit", "gp", "radon100k"):
        chains = int(min(args.chains, max(chains, chains * 15.0 / max(dt, 1e-3)) // cores * cores))
        evals, dt = cpu_sample_rate(ir, args, chains, args.seed)
        rate = evals / dt
    return {
        "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": "%d of the %d chains, %s, %s, OpenMP over chains (%d threads); %.3g grad-evals in %.1f s"
                  % (chains, args.chains,
                     {"logit": "2 + 2 transitions", "gp": "one gradient evaluation each",
                      "radon100k": "10 + 2 transitions"}.get(WORKLOAD, "same burn_in/draws"),
                     "float64" if WORKLOAD == "gp" else "float32", cores, evals, dt),
    }


def tfp_reference_rate(args):
    """when TensorFlow Probability is importable on this box (it is not in this image: no wheel in
    /opt/wheelhouse, no network), time the sampler the reference really runs -- tfp.mcmc.NoUTurnSampler under
    DualAveragingStepSizeAdaptation, vectorised over chains, CPU only -- on the same synthetic radon inputs.  Returns
    (grad-evals, seconds, description) or None."""
    if WORKLOAD != "radon919":
        return None
    try:
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
        import tensorflow as tf
        import tensorflow_probability as tfp
    except Exception:
        return None
    from pymc4_b200 import benchmodels as M

    idx, x, y, G = M.radon_synthetic(N_OBS, N_GROUPS)
    idx_t, x_t, y_t = tf.constant(idx), tf.constant(x, tf.float32), tf.constant(y, tf.float32)
    tfd = tfp.distributions

    def logp(mu_a, mu_b, a, b, ls_a, ls_b, l_eps):  # the reference's joint density in unconstrained space
        sa, sb, eps = tf.exp(ls_a), tf.exp(ls_b), tf.exp(l_eps)
        lp = tfd.Normal(0., 1.).log_prob(mu_a) + tfd.Normal(0., 1.).log_prob(mu_b)
        lp += tfd.HalfCauchy(0., 1.).log_prob(sa) + ls_a + tfd.HalfCauchy(0., 1.).log_prob(sb) + ls_b
        lp += tfd.HalfCauchy(0., 1.).log_prob(eps) + l_eps
        lp += tf.reduce_sum(tfd.Normal(mu_a[..., None], sa[..., None]).log_prob(a), -1)
        lp += tf.reduce_sum(tfd.Normal(mu_b[..., None], sb[..., None]).log_prob(b), -1)
        est = tf.gather(a, idx_t, axis=-1) + tf.gather(b, idx_t, axis=-1) * x_t
        return lp + tf.reduce_sum(tfd.Normal(est, eps[..., None]).log_prob(y_t), -1)

    C = args.cpu_chains or 64
    init = [tf.zeros([C]), tf.zeros([C]), tf.zeros([C, G]), tf.zeros([C, G]), tf.zeros([C]), tf.zeros([C]), tf.zeros([C])]
    kernel = tfp.mcmc.DualAveragingStepSizeAdaptation(
        tfp.mcmc.NoUTurnSampler(logp, step_size=args.step_size), num_adaptation_steps=args.burn_in,
        step_size_setter_fn=lambda pkr, s: pkr._replace(step_size=s), step_size_getter_fn=lambda pkr: pkr.step_size,
        log_accept_prob_getter_fn=lambda pkr: pkr.log_accept_ratio)

    @tf.function(autograph=False)
    def run():
        return tfp.mcmc.sample_chain(args.draws, init, kernel=kernel, num_burnin_steps=args.burn_in,
                                     trace_fn=lambda _, pkr: pkr.inner_results.leapfrogs_taken, seed=args.seed)

    run()  # trace + warm-up
    t0 = time.perf_counter()
    _, leaps = run()
    dt = time.perf_counter() - t0
    # burn-in leapfrogs are not traced by sample_chain: scale the kept draws' count to all transitions
    evals = float(tf.reduce_sum(leaps)) * (args.burn_in + args.draws) / max(args.draws, 1)
    return evals, dt, "tfp.mcmc.NoUTurnSampler + DualAveragingStepSizeAdaptation (TF %s, TFP %s), %d chains, CPU" % (
        tf.__version__, tfp.__version__, C)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tfp_run = tfp_reference_rate(args)
    if tfp_run is not None:
        evals, secs, what = tfp_run
        value = evals / secs
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": 1,
            "warmup": 1, "ms_per_step": 1e3 * secs, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference", "sample": what},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    _, ir = build_model()
    from oracle import pm4_oracle

    pm4_oracle.build()
    cores = pm4_oracle.set_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: use every host thread
    chains = args.cpu_chains or (cores if WORKLOAD in ("logit", "gp") else 2 * cores)
    for _ in range((max(args.warmup, 0) and 1) if WORKLOAD not in ("logit", "gp") else 0):  # one short warm-up pass is enough for a CPU code
        cpu_sample_rate(ir, args, cores, args.seed + 1)
    evals, secs = 0.0, 0.0
    for k in range(args.steps):
        e, dt = cpu_sample_rate(ir, args, chains, args.seed + k)
        evals += e
        secs += dt
    value = evals / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if WORKLOAD == "gp" else "f32",
        "data": "synthetic", "config": workload_config(args, args.gpus),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "each step = %d of the %d chains per GPU, %s; TFP-semantics CPU oracle "
                      "(the reference's TF/TFP stack is not installable in this image), OpenMP over chains" %
                      (chains, args.chains, {"logit": "2 + 2 transitions", "gp": "one gradient evaluation each",
                                        "radon100k": "10 + 2 transitions"}.get(WORKLOAD, "same burn_in/draws")),
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_ready(self, timeout=5.0):
        """Block until the first sample has arrived (the start-up of nvidia-smi is over)."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout and self.proc.poll() is None:
            time.sleep(0.02)

    def mark(self):
        """The timed region starts now: earlier samples are dropped."""
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[getattr(self, "first", 0):]:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def onchip_peaks(torch):
    """FP32-FMA and shared-memory peaks measured now on this GPU (csrc/pm4_peaks.cu)."""
    import ctypes

    from pymc4_b200 import codegen

    src = os.path.join(ROOT, "pymc4_b200", "csrc", "pm4_peaks.cu")
    lib = os.path.join(ROOT, "pymc4_b200", "libpm4peaks.so")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.run([codegen.find_nvcc()] + codegen.NVCC_FLAGS + ["-shared", "-o", lib, src], check=True)
    L = ctypes.CDLL(lib)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    fp32, smem = ctypes.c_double(0), ctypes.c_double(0)
    rc1 = L.pm4_peak_fp32_tflops(2, 20000, st, ctypes.byref(fp32))
    rc2 = L.pm4_peak_smem_tbs(20000, st, ctypes.byref(smem))
    return (fp32.value if rc1 == 0 else None), (smem.value if rc2 == 0 else None)


# ------------------------------------------------------------------------------------------
# the other BASELINE configurations, bounded, as the default line's "secondary" object
# ------------------------------------------------------------------------------------------
SECONDARY = {
    # configs[2]: hierarchical logistic regression, N = 1M x P = 100, 1024 chains per GPU (HMC, 3 leapfrogs)
    "logit": ["--workload", "logit", "--steps", "2", "--warmup", "1"],
    # configs[3]: marginal GP regression, n = 4096, 256 chains per GPU (HMC, 3 leapfrogs); one batched gradient
    # evaluation of 256 chains is 1.8e13 float64 flops, so the run is 1 + 1 transitions
    "gp": ["--workload", "gp", "--burn-in", "1", "--draws", "1", "--steps", "1", "--warmup", "1"],
    # configs[4]: radon-shaped, 100k observations / 5k groups, 8192 chains per GPU (65 536 chains at 8 GPUs)
    "radon100k": ["--workload", "radon100k", "--chains", "8192", "--burn-in", "20", "--draws", "10", "--steps", "1",
                  "--warmup", "1"],
}


def run_secondaries(args, rank, local_rank):
    """Every rank runs the bounded secondary workloads as single-GPU child processes on ITS GPU (shard `rank` of the
    box-wide run: global chain offsets, no data-path collective; the scalar step size is pooled over the rank's
    chains).  Returns {name: parsed JSON line or {"error": ...}}."""
    env = dict(os.environ)
    for k in list(env):
        if k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK", "MASTER_ADDR",
                 "MASTER_PORT") or k.startswith("TORCHELASTIC"):
            env.pop(k)
    visible = [v for v in env.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
    env["CUDA_VISIBLE_DEVICES"] = visible[local_rank] if local_rank < len(visible) else str(local_rank)
    env.pop("PM4_PHASE_TIMES", None)
    out = {}
    for name, argv in SECONDARY.items():
        cmd = [sys.executable, os.path.abspath(__file__), "--gpus", "1", "--no-cpu-baseline", "--no-secondary",
               "--shard-rank", str(rank), "--seed", str(args.seed)] + argv
        t0 = time.perf_counter()
        try:
            proc = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
            lines = [ln for ln in proc.stdout.strip().splitlines() if ln.startswith("{")]
            if proc.returncode != 0 or not lines:
                out[name] = {"error": "exit %d: %s" % (proc.returncode, proc.stderr.strip()[-300:])}
            else:
                out[name] = json.loads(lines[-1])
                out[name]["child_wall_s"] = time.perf_counter() - t0
        except subprocess.TimeoutExpired:
            out[name] = {"error": "timed out after 600 s"}
    return out


def merge_secondaries(per_rank):
    """Whole-job figures of the secondary workloads: units of all ranks over the slowest rank's device time."""
    merged = {}
    for name in SECONDARY:
        rows = [r.get(name) for r in per_rank if r]
        bad = [r for r in rows if not r or "error" in r]
        if bad:
            merged[name] = {"error": (bad[0] or {}).get("error", "missing")}
            continue
        secs = [r["ms_per_step"] * r["steps"] / 1e3 for r in rows]
        units = sum(r["value"] * t for r, t in zip(rows, secs))
        e_secs = [r["e2e"]["ms_per_step"] / 1e3 for r in rows if r.get("e2e")]
        e_units = sum(r["e2e"]["value"] * t for r, t in zip(rows, e_secs)) if e_secs else None
        first = rows[0]
        merged[name] = {
            "metric": first["metric"], "value": units / max(secs), "unit": first["unit"], "n_gpus": len(rows),
            "ms_per_step": 1e3 * max(secs) / first["steps"], "steps": first["steps"], "warmup": first["warmup"],
            "dtype": first["dtype"], "scaling": "weak", "config": first["config"],
            "parallelism": "one single-GPU process per rank on its own GPU (chains sharded by global offset, every GPU "
                           "holds the data, no data-path collective; scalar step size pooled over the rank's chains)",
            "e2e": ({"value": e_units / max(e_secs), "unit": first["unit"],
                     "h2d_bytes_per_step": first["e2e"]["h2d_bytes_per_step"],
                     "d2h_bytes_per_step": first["e2e"]["d2h_bytes_per_step"], "api": first["e2e"].get("api")}
                    if e_secs else None),
            "roofline": first["roofline"], "gpu_launches": first.get("gpu_launches"), "clocks": first.get("clocks"),
            "ess_per_sec": first.get("ess_per_sec"), "diagnostics": first.get("diagnostics"),
            "per_rank_values": [r["value"] for r in rows],
        }
    return merged


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sampler has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from pymc4_b200 import _cabi, diagnostics
    from pymc4_b200._cstructs import PM4_EPS_PER_CHAIN, PM4_EPS_POOLED, make_adapt, make_nuts_cfg
    from pymc4_b200.mcmc.samplers import NUTS

    model, ir = build_model()
    dm = _cabi.DeviceModel(ir, device=local_rank)
    C, B, S = args.chains, args.burn_in, args.draws
    shard = args.shard_rank if (world == 1 and args.shard_rank >= 0) else rank  # global chain offsets of this process
    comm = None
    if world > 1 and args.eps_mode == "pooled" and args.couple_step_size:
        # one pm.sample call sharded over the box: the scalar step size is adapted from the mean acceptance
        # of the chains of ALL ranks (mailboxes in peer memory, read over NVLink by the dual-averaging kernel)
        comm = _cabi.Communicator(local_rank, rank, world)
        dm.attach_comm(comm)
    adapt = make_adapt(mode=PM4_EPS_PER_CHAIN if args.eps_mode == "per-chain" else PM4_EPS_POOLED,
                       num_adaptation_steps=B, mass_windows=args.mass_windows)

    is_logit = args.workload in ("logit", "gp")  # the lock-step HMC workloads

    def cfg_for(step):
        if is_logit:
            from pymc4_b200._cstructs import make_hmc_cfg

            return make_hmc_cfg(C, S, B, step_size=args.step_size, num_leapfrog_steps=args.leapfrogs,
                                seed=args.seed + step, adapt=adapt, chain_offset=shard * C)
        return make_nuts_cfg(C, S, B, step_size=args.step_size, seed=args.seed + step, adapt=adapt,
                             max_tree_depth=args.max_tree_depth,
                             warps_per_block=args.warps_per_block, chain_offset=shard * C)

    def run_step(cfg):
        if not is_logit:
            return dm.nuts(cfg, theta0, dtype=np.float32)
        o = dm.hmc(cfg, theta0, dtype=np.float32)
        # every chain takes exactly L gradient evaluations per transition
        o["total_leapfrogs"] = torch.full((C,), (B + S) * args.leapfrogs, dtype=torch.int64, device=dev)
        o["leapfrogs"] = torch.full((S, C), args.leapfrogs, dtype=torch.int32, device=dev)
        o["diverging"] = torch.zeros((S, C), dtype=torch.uint8, device=dev)
        o["log_accept"] = o["log_accept"].clamp(max=0.0)
        return o

    theta0 = torch.zeros((C, ir.D), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: inputs already in HBM ------------------------------------
    # the clock sampler (one long-running `nvidia-smi -lms 100`) starts BEFORE the warm-up steps: its start-up (NVML
    # attach, ~0.5 s) stalls whatever step it lands in by ~100 ms; only the samples taken during the timed region count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    for w in range(args.warmup):
        out = run_step(cfg_for(1000 + w))
    torch.cuda.synchronize()
    if rank == 0:
        sampler.mark()
    barrier()
    _cabi.launch_count(reset=True)
    step_ms, leaps, detail = [], 0, []
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # evict L2 between timed steps (not timed)
        if comm is not None:
            barrier()  # coupled ranks start a step together (a rank waits for its peers inside the step)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = run_step(cfg_for(k))
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        per_chain = out["total_leapfrogs"]
        leaps += int(per_chain.sum().item())
        # a chain is sequential: the longest one bounds the step from below (critical path)
        detail.append({"ms": step_ms[-1], "leapfrogs": int(per_chain.sum().item()),
                       "max_chain_leapfrogs": int(per_chain.max().item()),
                       "mean_chain_leapfrogs": float(per_chain.float().mean().item())})
    barrier()
    wall = time.perf_counter() - wall0
    launches = _cabi.launch_count()
    clocks = sampler.stop() if rank == 0 else None

    dev_secs = sum(step_ms) / 1e3
    t = torch.tensor([dev_secs, float(leaps)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_secs_max, total_leaps = float(tmax[0]), float(tsum[1])
    else:
        dev_secs_max, total_leaps = dev_secs, float(leaps)
    value = total_leaps / dev_secs_max

    # ---- diagnostics on the last step (all-gather of the monitored columns over NCCL) ---
    cols = [0, 1, 2, 2 + N_GROUPS, ir.D - 3, ir.D - 2, ir.D - 1]  # mu_a, mu_b, alpha[0], beta[0], log sigmas, log eps
    if is_logit:
        cols = [0, 1, 2, ir.D // 2, ir.D - 2, ir.D - 1]  # mu, beta[0], beta[1], beta[mid], beta[last], log tau
    if args.workload == "gp":
        cols = [0, 1, 2]  # log l, log a, log sigma
    summ = diagnostics.summarize(out["states"], columns=cols)
    ess_per_sec = summ["min_ess"] / (dev_secs_max / max(args.steps, 1))
    mean_accept = float(torch.exp(out["log_accept"]).mean().item())
    divergent = float(out["diverging"].float().mean().item())
    stuck_frac = float((out["diverging"].float().mean(dim=0) > 0.9).float().mean().item())  # chains diverging on ~every transition
    mean_tree = float(out["leapfrogs"].float().mean().item())

    # ---- the same workload with diagonal mass-matrix windows (one extra step, reported beside the headline) -------
    mass_adapted = None
    if args.workload == "radon919" and args.mass_windows == 0 and not is_logit:
        adapt_m = make_adapt(mode=PM4_EPS_PER_CHAIN if args.eps_mode == "per-chain" else PM4_EPS_POOLED,
                             num_adaptation_steps=B, mass_windows=4)
        cfg_m = make_nuts_cfg(C, S, B, step_size=args.step_size, seed=args.seed + 77, adapt=adapt_m,
                              max_tree_depth=args.max_tree_depth, warps_per_block=args.warps_per_block, chain_offset=shard * C)
        dm.nuts(cfg_m, theta0, dtype=np.float32)  # warm-up of the scaled kernel variant
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        om = dm.nuts(cfg_m, theta0, dtype=np.float32)
        e1.record()
        e1.synchronize()
        ms_m = e0.elapsed_time(e1)
        sm = diagnostics.summarize(om["states"], columns=cols)
        lm = float(om["total_leapfrogs"].sum().item())
        dvm = om["diverging"].float()
        mass_adapted = {
            "what": "same workload, one step, 4 diagonal mass-matrix windows during burn-in (pm.sample(..., mass_windows=4); "
                    "not in the reference API: its radon notebook hand-feeds per-dimension step sizes instead)",
            "value_this_rank": lm / (ms_m / 1e3), "unit": UNIT, "ms_per_step": ms_m,
            "ess_per_sec_this_rank": sm["min_ess"] / (ms_m / 1e3),
            "diagnostics": {"min_bulk_ess": sm["min_ess"], "max_split_rhat": sm["max_rhat"],
                            "divergent_frac": float(dvm.mean().item()),
                            "stuck_chain_frac": float((dvm.mean(dim=0) > 0.9).float().mean().item()),
                            "mean_leapfrogs_per_draw": float(om["leapfrogs"].float().mean().item()),
                            "step_size": float(om["step_size"][0].item())},
        }
        del om

    # ---- the same data with the group effects written as offsets (one extra step): the configuration on which the
    # reference-default sampler settings meet north_star check (c) -- split R-hat < 1.01, no divergences to speak of
    noncentered = None
    if args.workload == "radon919" and args.mass_windows == 0 and not is_logit and N_OBS == 919:
        from pymc4_b200 import benchmodels as BM
        from pymc4_b200.ir import lower_model as _lower

        ir_n, _, _ = _lower(BM.hierarchical_model_noncentered(*BM.radon_synthetic(N_OBS, N_GROUPS)))
        dm_n = _cabi.DeviceModel(ir_n, device=local_rank)
        cfg_n = make_nuts_cfg(C, S, B, step_size=args.step_size, seed=args.seed + 78, adapt=adapt,
                              max_tree_depth=args.max_tree_depth, warps_per_block=args.warps_per_block, chain_offset=shard * C)
        dm_n.nuts(cfg_n, theta0, dtype=np.float32)  # warm-up
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        on = dm_n.nuts(cfg_n, theta0, dtype=np.float32)
        e1.record()
        e1.synchronize()
        ms_n = e0.elapsed_time(e1)
        sn = diagnostics.summarize(on["states"], columns=cols)
        noncentered = {
            "what": "same data, same sampler settings, group effects written as offsets (alpha = mu_alpha + sigma_alpha * "
                    "alpha_raw; pymc4_b200.benchmodels.hierarchical_model_noncentered): one step",
            "value_this_rank": float(on["total_leapfrogs"].sum().item()) / (ms_n / 1e3), "unit": UNIT, "ms_per_step": ms_n,
            "ess_per_sec_this_rank": sn["min_ess"] / (ms_n / 1e3),
            "diagnostics": {"min_bulk_ess": sn["min_ess"], "max_split_rhat": sn["max_rhat"],
                            "divergent_frac": float(on["diverging"].float().mean().item()),
                            "mean_leapfrogs_per_draw": float(on["leapfrogs"].float().mean().item()),
                            "step_size": float(on["step_size"][0].item())},
        }
        del on
        dm_n.close()

    # ---- end-to-end arm: public sampler API, host buffers in, host trace out --------------
    e2e = None
    if not args.no_e2e:
        step_arr = np.full((C, 1), args.step_size, dtype=np.float32) if args.eps_mode == "per-chain" else args.step_size
        extra = {"comm": comm} if comm is not None else {}
        if is_logit:
            from pymc4_b200.mcmc.samplers import HMC

            nuts = HMC(model, step_size=step_arr, num_leapfrog_steps=args.leapfrogs, num_adaptation_steps=B,
                       device=local_rank, chain_offset=shard * C, **extra)
        else:
            nuts = NUTS(model, step_size=step_arr, num_adaptation_steps=B, device=local_rank, chain_offset=shard * C,
                        **extra)
        e2e_leaps, e2e_secs, h2d, d2h = 0.0, 0.0, 0, 0
        e2e_pass_ms = []
        n_e2e = max(1, min(args.steps, 2))
        n_e2e_warm = 2  # untimed passes: allocator pools, pinned trace buffer, lazily sized workspaces
        for k in range(n_e2e_warm + n_e2e):
            trace = None  # the caller drops the previous trace (its pinned host blocks go back to torch's host allocator)
            gc.collect()
            barrier()
            t0 = time.perf_counter()
            trace = nuts(num_samples=S, num_chains=C, burn_in=B, seed=args.seed + 500 + k)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            e2e_pass_ms.append(1e3 * dt)
            if k < n_e2e_warm:
                continue
            e2e_secs += dt
            e2e_leaps += (float(C * (B + S) * args.leapfrogs) if is_logit
                          else float(nuts.last_run_info["total_leapfrogs"].sum().item()))
            h2d = C * ir.D * 4 + ir.dataf.size * 4 + ir.datai.size * 4
            d2h = sum(v.nbytes for v in trace.posterior.values()) + sum(v.nbytes for v in trace.sample_stats.values())
        te = torch.tensor([e2e_secs, e2e_leaps], dtype=torch.float64, device=dev)
        if world > 1:
            a = te.clone()
            dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = te.clone()
            dist.all_reduce(b, op=dist.ReduceOp.SUM)
            e2e_secs, e2e_leaps = float(a[0]), float(b[1])
        e2e = {"value": e2e_leaps / e2e_secs, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_secs / n_e2e,
               "passes_ms": [round(v, 1) for v in e2e_pass_ms],  # the first two passes (allocator / pinned-buffer warm-up) are not timed
               "api": "pymc4_b200.mcmc.samplers.%s(model, step_size=%s)(num_samples, num_chains, burn_in) -> InferenceData "
                      "on the host (trace copied through pinned memory)"
                      % ("HMC" if is_logit else "NUTS", "[C,1]" if args.eps_mode == "per-chain" else "%g" % args.step_size)}

    secondary = None
    if args.workload == "radon919" and not args.no_secondary:
        mine = run_secondaries(args, rank, local_rank)
        gathered = [mine]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
        secondary = merge_secondaries(gathered)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peaks_src = measured_peaks()
    fp32_peak, smem_peak = onchip_peaks(torch)
    units_per_step_gpu = leaps / max(args.steps, 1)
    kernel_secs = float(np.mean(step_ms)) / 1e3
    ach_tflops = units_per_step_gpu * FLOPS_PER_UNIT / kernel_secs / 1e12
    ach_smem_tbs = units_per_step_gpu * ONCHIP_BYTES_PER_UNIT / kernel_secs / 1e12
    # for the 919-observation shape state and data are on-chip; of the two on-chip roofs the
    # shared-memory read roof (12 bytes per observation per unit) is the tighter one and is "the roofline for
    # its log-prob pass"; the FP32-pipe figure is reported beside it, HBM for context
    roofline = {
        "bound": "smem", "achieved": ach_smem_tbs * 1e3, "peak": (smem_peak * 1e3) if smem_peak else None, "unit": "GB/s",
        "frac": (ach_smem_tbs / smem_peak) if smem_peak else None,
        # measured shared-memory traffic per unit of the dominant kernel, from the committed ncu --set full capture
        # (profiles/r02_nuts_warp_kernel_summary.md): l1tex__data_pipe_lsu_wavefronts_mem_shared.sum x 128 B / leapfrogs of
        # the launch = 167 wavefronts = 21.4 KB per unit against 11.0 KB algorithmic (records are read as 16-byte pair
        # rows, the tree checkpoints and the subtree candidate live in shared memory); DRAM traffic is the trace only
        "traffic": (21.4e3 * units_per_step_gpu) if WORKLOAD == "radon919" else None,
        "traffic_unit": "shared-memory wavefront bytes per step (167 wavefronts x 128 B per unit, ncu)",
        "kernel": ("pm4::nuts_warp_kernel (one warp per chain, owner-computes layout; per step: ONE cooperative launch for the "
                   "burn_in + 1 lock-step transitions while the pooled step size adapts -- grid barrier + dual-averaging "
                   "update inside the kernel -- then one persistent launch for the draws; bootstrap and export kernels are "
                   "<0.1% of the step)") if WORKLOAD == "radon919" and args.eps_mode == "pooled" else
                  "pm4::nuts_warp_kernel / pm4::nuts_kernel (persistent blocks; bootstrap, chain-ordering and export "
                  "kernels are <0.1% of the step)",
        "algorithmic_bytes_per_unit": ONCHIP_BYTES_PER_UNIT,
        "peak_source": "LDS.128 microbenchmark run by this bench on this GPU (csrc/pm4_peaks.cu); "
                       "MEASURED_PEAKS.json has no shared-memory entry",
        "fp32": {"achieved": ach_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                 "frac": (ach_tflops / fp32_peak) if fp32_peak else None,
                 "algorithmic_flops_per_unit": FLOPS_PER_UNIT,
                 "peak_source": "FFMA microbenchmark run by this bench on this GPU (csrc/pm4_peaks.cu)"},
        "hbm": {"peak_gbs": peaks.get("hbm_gbs"), "source": peaks_src,
                "note": "state and data are on-chip for this shape; HBM carries only the trace writes"},
    }
    if args.workload == "radon100k":
        # for the 100k shape the headline fraction is against HBM with 24 D bytes per unit
        # (theta, momentum, gradient read and written once); this build keeps them in registers, so the
        # figure says how far the kernel is from the point where HBM would start to matter
        hbm_peak = peaks.get("hbm_gbs")
        ach = units_per_step_gpu * 24.0 * D_MODEL / kernel_secs / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": (ach / hbm_peak) if hbm_peak else None, "traffic": None,
                    "kernel": roofline["kernel"], "algorithmic_bytes_per_unit": 24 * D_MODEL,
                    "fp32": {"achieved": ach_tflops, "peak": fp32_peak, "unit": "TFLOP/s"},
                    "l2_records": {"achieved": ach_smem_tbs, "unit": "TB/s",
                                   "note": "observation records (12 N bytes per unit) stream from L2, not shared memory"}}
    if args.workload == "gp":
        # ~n^3 float64 flops per unit; yardstick = cuBLAS DGEMM timed live on this GPU
        ma = torch.randn(4096, 4096, device=dev, dtype=torch.float64)
        mb = torch.randn(4096, 4096, device=dev, dtype=torch.float64)
        for _ in range(2):
            ma @ mb
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ma @ mb
        e1.record()
        torch.cuda.synchronize()
        f64_peak = 10 * 2 * 4096.0 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        roofline = {
            "bound": "tensor", "achieved": ach_tflops, "peak": f64_peak, "unit": "TFLOP/s",
            "frac": ach_tflops / f64_peak, "traffic": None,
            "kernel": "gp_nt_kernel<CHOL|ZT|DOT> (64 x 64 float64 tiles on the FP64 tensor cores, DMMA m8n8k4): blocked Cholesky, inverse factor and the "
                      "contraction with dK; time base = the whole step",
            "algorithmic_flops_per_unit": FLOPS_PER_UNIT,
            "peak_source": "cuBLAS float64 matmul 4096^3 timed live by this bench (MEASURED_PEAKS.json has no float64 entry)",
        }
    elif is_logit:
        # the dense likelihood is GEMM-shaped: 4 N P useful flops per gradient evaluation (X beta and
        # X^T r).  The kernels issue them as 3xTF32 (three tensor-core passes per product, TF32 at half the bf16
        # rate), so the fraction against the measured bf16 peak is bounded by 1/6 for FP32-accurate results.
        tpeak = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops")
        torch.backends.cuda.matmul.allow_tf32 = True
        ma = torch.randn(8192, 8192, device=dev)
        mb = torch.randn(8192, 8192, device=dev)
        for _ in range(2):
            ma @ mb
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ma @ mb
        e1.record()
        torch.cuda.synchronize()
        tf32_peak = 10 * 2 * 8192.0 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        roofline = {
            "bound": "tensor", "achieved": ach_tflops, "peak": tpeak, "unit": "TFLOP/s",
            "frac": (ach_tflops / tpeak) if tpeak else None, "traffic": None,
            "kernel": "glm_fused_kernel (warp-specialised: TMA bulk copies of the pre-split X tile images, tcgen05.mma kind::tf32 "
                      "with both A operands in tensor memory, Bernoulli epilogue between the two GEMMs; the residual matrix "
                      "never leaves the SM), 1 launch + 1 reduction per lock-step gradient evaluation of all chains; "
                      "time base = the whole step",
            "algorithmic_flops_per_unit": FLOPS_PER_UNIT,
            "peak_source": peaks_src + " bf16 sustained",
            "tf32": {"issued_tflops": 3.0 * ach_tflops, "cublas_tf32_peak_tflops": tf32_peak,
                     "frac_of_tf32_issued": 3.0 * ach_tflops / tf32_peak,
                     "note": "3xTF32 split: hi*hi + lo*hi + hi*lo keeps ~22 mantissa bits; cuBLAS TF32 8192^3 timed live"},
            "hbm": {"achieved_gbs": units_per_step_gpu / C * 4.0 * 4.0 * N_OBS * LOGIT_P / kernel_secs / 1e9,
                    "peak_gbs": peaks.get("hbm_gbs"),
                    "note": "per lock-step gradient evaluation of all chains the pre-split image (hi and lo TF32 terms of X "
                            "and of its transposed tiles: 4 x the bytes of X) is read once from HBM; the chain tiles that "
                            "share an observation tile find it in L2 (ncu: 1.73 GB per evaluation)"},
        }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dev_secs_max / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if WORKLOAD == "gp" else "f32",
        "data": "synthetic", "config": workload_config(args, world),
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "ess_per_sec": ess_per_sec,
        "diagnostics": {"min_bulk_ess": summ["min_ess"], "max_split_rhat": summ["max_rhat"],
                        "columns": cols, "chains": summ["n_chains"], "draws": summ["n_draws"],
                        "mean_accept": mean_accept, "divergent_frac": divergent,
                        "mean_leapfrogs_per_draw": mean_tree},
        "wall_s_timed_region": wall,
        "steps_detail": detail,
    }
    if mass_adapted is not None:
        line["mass_adapted"] = mass_adapted
    if noncentered is not None:
        line["noncentered"] = noncentered
    line["diagnostics"]["stuck_chain_frac"] = stuck_frac
    if secondary is not None:
        line["secondary"] = secondary
    if not args.no_cpu_baseline:  # rank 0's host cores, a bounded sample of the per-GPU workload (also on N > 1 lines)
        line["cpu_baseline"] = cpu_baseline(ir, args)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    select_workload(args.workload, args.logit_n, args.gp_n)
    if args.workload == "gp":
        # one batched gradient evaluation of 256 chains is 1.8e13 float64 flops: bounded defaults
        if args.chains == CHAINS_PER_GPU:
            args.chains = 256
        if args.burn_in == 1000 and args.draws == 1000:
            args.burn_in, args.draws = 2, 2
        if args.step_size == 0.1:
            args.step_size = 0.005
    if args.workload == "logit":
        # bounded defaults for the dense workload (one batched leapfrog of 1024 chains reads the 400 MB matrix twice)
        if args.chains == CHAINS_PER_GPU:
            args.chains = 1024
        if args.burn_in == 1000 and args.draws == 1000:
            args.burn_in, args.draws = 10, 10
        if args.step_size == 0.1:
            args.step_size = 0.002
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
