"""CPU restatement of the reference's posteriorgram post-processing — the step right after
the hot path.  TEST INFRASTRUCTURE ONLY: imported by `tests/` to check the
CUDA kernels of `ppgs_b200/csrc/postops.cu`; never by the product package.

Pinned against the reference's own functions run in the dev container
(`oracle/make_golden_postops.py` -> `tests/golden/postops_s*.npz`, and live in
`tests/test_postops_oracle.py` when /root/reference exists).  The reference ships no tests for
these functions either."""
import torch

SIMILARITY_EXPONENT = 1.2   # ppgs/config/defaults.py:214


def distance(ppgX, ppgY, reduction='mean', normalize=True, exponent=SIMILARITY_EXPONENT,
             similarity=None):
    """ppgs.distance (ppgs/core.py:399-469): similarity-weighted Jensen-Shannon distance of two
    aligned (phonemes, frames) posteriorgrams.  `similarity` replaces the matrix the
    reference loads from its assets (ppgs/core.py:436-442)."""
    ppgX = torch.clamp(ppgX, 1e-8, 1 - 1e-8)
    ppgY = torch.clamp(ppgY, 1e-8, 1 - 1e-8)
    if normalize:
        weights = similarity.to(ppgX.dtype).T ** exponent
        ppgX = (weights @ ppgX).T
        ppgY = (weights @ ppgY).T
    else:
        ppgX, ppgY = ppgX.T, ppgY.T
    log_average = torch.log((ppgX + ppgY) / 2)
    # torch.nn.functional.kl_div(input, target, reduction='none') evaluates
    # xlogy(target, target) - target * input; the square root below amplifies the rounding of
    # near-zero divergences, so the same association is kept
    kl_X = torch.xlogy(ppgX, ppgX) - ppgX * log_average
    kl_Y = torch.xlogy(ppgY, ppgY) - ppgY * log_average
    average_kl = torch.clamp((kl_X + kl_Y) / 2, min=0)
    jsd = torch.sqrt(average_kl).sum(dim=1)
    if reduction == 'mean':
        return jsd.mean(dim=0)
    if reduction == 'none' or reduction is None:
        return jsd
    if reduction == 'sum':
        return jsd.sum(dim=0)
    raise ValueError(f'Reduction method {reduction} not defined')


def interpolate(ppgX, ppgY, interp):
    """ppgs.interpolate (ppgs/core.py:475-496)."""
    return (1. - interp) * ppgX + interp * ppgY


def sparsify(ppg, method='percentile', threshold=0.85):
    """ppgs.sparsify (ppgs/core.py:504-543) for (batch, phonemes, frames).  'topk' follows
    the reference for batch = 1 (its advanced-indexing loop at :535-537 mixes the rows of a
    larger batch); every row is treated like that single row here."""
    if method in ('constant', 'percentile'):
        if method == 'percentile':
            q = torch.as_tensor([threshold], dtype=ppg.dtype).reshape(-1)
            threshold = torch.quantile(ppg, q, dim=-2, keepdim=True)[0]
        ppg = torch.where(ppg > threshold, ppg, torch.zeros_like(ppg))
    elif method == 'topk':
        values, indices = ppg.topk(int(threshold), dim=-2)
        ppg = torch.zeros_like(ppg).scatter(-2, indices, values)
    else:
        raise ValueError(f'Sparsification method {method} not defined')
    return torch.softmax(torch.log(ppg + 1e-8), -2)


def grid_sample(ppg, grid):
    """ppgs.edit.grid.sample (ppgs/edit/grid.py:13-50): linear interpolation of the frames at
    float-valued indices, final frame replicated."""
    interp = grid - torch.floor(grid)
    xp = torch.arange(ppg.shape[-1])
    i = torch.searchsorted(xp, grid, side='right')
    ppg = torch.nn.functional.pad(ppg, (0, 1), mode='replicate')
    return interpolate(ppg[..., i - 1], ppg[..., i], interp)


def random_similarity(seed=0, phonemes=40):
    """A seeded stand-in for assets/balanced_similarity.pt: positive, diagonally dominant."""
    g = torch.Generator().manual_seed(seed)
    matrix = torch.rand(phonemes, phonemes, generator=g) * 0.2 + torch.eye(phonemes)
    return matrix / matrix.sum(0, keepdim=True)


def random_ppg(seed, frames, batch=None, phonemes=40, sharpness=4.0):
    g = torch.Generator().manual_seed(seed)
    shape = (phonemes, frames) if batch is None else (batch, phonemes, frames)
    return torch.softmax(torch.randn(shape, generator=g) * sharpness, dim=-2)
