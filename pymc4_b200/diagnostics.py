"""Cross-chain convergence diagnostics on the device: split R-hat and bulk effective sample size.

Mirrors what the reference gets from ArviZ after ``trace_to_arviz`` (``az.summary`` in
/root/reference/notebooks/radon_hierarchical.ipynb): rank-normalised split-R-hat and bulk ESS
(Vehtari et al. 2021).  Inputs are device tensors laid out ``[draw, chain]`` as the sampler
writes them; everything stays on the GPU (torch is used for sort / FFT plumbing).  With chains
sharded over ranks, ``gather_chains`` all-gathers the selected columns over NCCL first -- the
only collective on the sampling path.
"""
from typing import Optional

import torch


def gather_chains(x: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather ``[draw, chain_local, ...]`` along the chain axis (no-op without a process group)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    parts = [torch.empty_like(x) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, x.contiguous(), group=group)
    return torch.cat(parts, dim=1)


def _split(x: torch.Tensor) -> torch.Tensor:
    """[draw, chain] -> [half, 2*chain]"""
    S = x.shape[0] // 2
    return torch.cat([x[:S], x[S: 2 * S]], dim=1)


def _rank_normalise(x: torch.Tensor) -> torch.Tensor:
    flat = x.reshape(-1)
    order = torch.argsort(flat)
    ranks = torch.empty_like(order, dtype=torch.float64)
    ranks[order] = torch.arange(1, flat.numel() + 1, device=x.device, dtype=torch.float64)
    p = (ranks - 0.375) / (flat.numel() + 0.25)
    z = torch.special.ndtri(p)
    return z.reshape(x.shape)


def split_rhat(x: torch.Tensor, rank_normalise: bool = True) -> float:
    """x: [draw, chain]"""
    x = x.to(torch.float64)
    z = _split(_rank_normalise(x) if rank_normalise else x)
    n = z.shape[0]
    w = z.var(dim=0, unbiased=True).mean()
    b = n * z.mean(dim=0).var(unbiased=True)
    return float(torch.sqrt(((n - 1) / n * w + b / n) / w))


def ess_bulk(x: torch.Tensor, rank_normalise: bool = True) -> float:
    """Bulk ESS of one scalar quantity, x: [draw, chain] (all chains pooled)."""
    x = x.to(torch.float64)
    z = _split(_rank_normalise(x) if rank_normalise else x)
    n, m = z.shape
    zc = z - z.mean(dim=0, keepdim=True)
    nfft = 1 << (2 * n - 1).bit_length()
    f = torch.fft.rfft(zc, n=nfft, dim=0)
    acov = torch.fft.irfft(f * f.conj(), n=nfft, dim=0)[:n] / n  # [lag, chain], biased estimator
    chain_var = acov[0] * n / (n - 1)
    w = chain_var.mean()
    var_plus = w * (n - 1) / n
    if m > 1:
        var_plus = var_plus + z.mean(dim=0).var(unbiased=True)
    rho = 1.0 - (w - acov.mean(dim=1)) / var_plus
    rho[0] = 1.0
    # Geyer: sums of adjacent pairs, truncated at the first negative pair, made monotone
    npair = n // 2
    pairs = rho[: 2 * npair].reshape(npair, 2).sum(dim=1)
    neg = (pairs < 0).nonzero()
    k = int(neg[0]) if neg.numel() else npair
    if k == 0:
        return float(n * m)
    pairs = torch.cummin(pairs[:k], dim=0).values
    tau = -1.0 + 2.0 * pairs.sum()
    tau = max(float(tau), 1.0 / torch.log10(torch.tensor(float(n * m))).item())
    return float(n * m / tau)


def summarize(states: torch.Tensor, columns: Optional[list] = None, group=None, gather: bool = True) -> dict:
    """min bulk-ESS and max split-R-hat over the selected columns of ``states`` [draw, chain, D].
    With ``gather`` (default) the chains of all ranks are pooled first: a collective every rank must call."""
    cols = list(range(states.shape[2])) if columns is None else list(columns)
    sel = states[:, :, cols].contiguous()
    if gather:
        sel = gather_chains(sel, group=group)
    if sel.shape[0] < 4:  # split chains need two draws per half
        nan = float("nan")
        return dict(min_ess=nan, max_rhat=nan, columns=cols, ess=[nan] * len(cols), rhat=[nan] * len(cols),
                    n_draws=int(sel.shape[0]), n_chains=int(sel.shape[1]))
    ess = [ess_bulk(sel[:, :, k]) for k in range(len(cols))]
    rhat = [split_rhat(sel[:, :, k]) for k in range(len(cols))]
    return dict(min_ess=min(ess), max_rhat=max(rhat), columns=cols, ess=ess, rhat=rhat,
                n_draws=int(sel.shape[0]), n_chains=int(sel.shape[1]))
