"""Generate the committed golden fixtures for the pm.sample hot path.

Run in the BUILD container (needs /root/reference, scipy):
    python tests/golden/make_golden.py

Outputs (committed):
  radon_919.npz      county_code / floor / log_radon columns of the reference's
                     notebooks/data/radon.csv (919 rows, 85 counties, rows sorted by county)
  golden_logp.json   * eight-schools log-density at eta=0, mu=0, __log_tau=1: the value the
                       reference asserts (tests/test_8schools.py:61) and its float64 restatement
                       with scipy.stats;
                     * radon hierarchical model (notebooks/radon_hierarchical.ipynb:72-95)
                       log-density on the real data at theta = 0 and at a seeded random point,
                       computed here with scipy.stats only (independent of this repo's code);
                     * the ExpQuad known-answer matrix of tests/test_gp_cov.py:10-20.
Nothing here imports pymc4_b200 or the oracle.
"""
import json
import os

import numpy as np
import pandas as pd
from scipy import stats

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def eight_schools_logp(eta, mu, log_tau):
    y = np.array([28, 8, -3, 7, -1, 1, 18, 12], dtype=np.float64)
    sigma = np.array([15, 10, 16, 11, 9, 11, 10, 18], dtype=np.float64)
    tau = np.exp(log_tau)
    lp = stats.norm.logpdf(eta, 0, 1).sum()
    lp += stats.norm.logpdf(mu, 1, 10)
    lp += stats.halfnorm.logpdf(tau, scale=2.0) + log_tau  # + log|d tau / d log_tau|
    lp += stats.norm.logpdf(y, mu + tau * eta, sigma).sum()
    return float(lp)


def radon_logp(theta, idx, x, y, G):
    # reference trace order: mu_alpha, mu_beta, alpha[G], beta[G], log sigma_alpha, log sigma_beta, log eps
    mu_a, mu_b = theta[0], theta[1]
    a, b = theta[2:2 + G], theta[2 + G:2 + 2 * G]
    lsa, lsb, le = theta[2 + 2 * G:]
    sa, sb, eps = np.exp(lsa), np.exp(lsb), np.exp(le)
    lp = stats.norm.logpdf(mu_a, 0, 1) + stats.norm.logpdf(mu_b, 0, 1)
    lp += stats.halfcauchy.logpdf(sa, scale=1) + lsa
    lp += stats.halfcauchy.logpdf(sb, scale=1) + lsb
    lp += stats.halfcauchy.logpdf(eps, scale=1) + le
    lp += stats.norm.logpdf(a, mu_a, sa).sum() + stats.norm.logpdf(b, mu_b, sb).sum()
    lp += stats.norm.logpdf(y, a[idx] + b[idx] * x, eps).sum()
    return float(lp)


def main():
    df = pd.read_csv(os.path.join(REF, "notebooks/data/radon.csv"))
    idx = df["county_code"].values.astype(np.int32)
    x = df["floor"].values.astype(np.float64)
    y = df["log_radon"].values.astype(np.float64)
    assert np.all(np.diff(idx) >= 0) and idx.max() == 84 and len(idx) == 919
    np.savez_compressed(os.path.join(HERE, "radon_919.npz"), idx=idx, x=x, y=y)

    G = 85
    D = 2 * G + 5
    theta_rand = np.random.default_rng(0).normal(0, 0.3, D)
    out = {
        "eight_schools": {
            "point": {"eta": [0.0] * 8, "mu": 0.0, "__log_tau": 1.0},
            "reference_assert_f32": -42.876114,  # /root/reference/tests/test_8schools.py:61
            "scipy_f64": eight_schools_logp(np.zeros(8), 0.0, 1.0),
        },
        "radon_919": {
            "order": "mu_alpha, mu_beta, alpha[85], beta[85], __log_sigma_alpha, __log_sigma_beta, __log_eps",
            "logp_at_zero": radon_logp(np.zeros(D), idx, x, y, G),
            "theta_rand": theta_rand.tolist(),
            "logp_at_rand": radon_logp(theta_rand, idx, x, y, G),
        },
        "expquad_kat": {  # /root/reference/tests/test_gp_cov.py:10-20
            "X": [[1.0, 2.0], [3.0, 4.0]], "amplitude": 1.0, "length_scale": 1.0,
            "K": [[1.0, 0.01831564], [0.01831564, 1.0]],
        },
    }
    with open(os.path.join(HERE, "golden_logp.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: (v if k != "radon_919" else {kk: vv for kk, vv in v.items() if kk != "theta_rand"})
                      for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main()
