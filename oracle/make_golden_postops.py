"""Golden vectors of the reference's post-processing functions (dev container only):
    python oracle/make_golden_postops.py
imports the UNMODIFIED reference through oracle/refshim.py, runs ppgs.distance /
ppgs.sparsify / ppgs.interpolate / ppgs.edit.grid.sample on seeded inputs and stores inputs
and outputs in tests/golden/postops_s<seed>.npz."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import postops_oracle as P  # noqa: E402
from oracle import refshim  # noqa: E402


def main():
    ppgs = refshim.import_reference()
    for seed, frames in ((0, 257), (1, 31)):
        x, y = P.random_ppg(seed, frames), P.random_ppg(seed + 100, frames, sharpness=1.0)
        similarity = P.random_similarity(seed)
        ppgs.distance.similarity_matrix = similarity          # the cache of ppgs/core.py:436-442
        ppgs.distance.device = x.device
        out = {'x': x, 'y': y, 'similarity': similarity}
        for reduction in ('mean', 'sum', 'none'):
            out[f'distance_{reduction}'] = ppgs.distance(x, y, reduction=reduction)
            out[f'distance_raw_{reduction}'] = ppgs.distance(x, y, reduction=reduction, normalize=False)
        out['distance_exp2'] = ppgs.distance(x, y, exponent=2.0)
        batch = P.random_ppg(seed + 7, frames, batch=1)
        out['batch'] = batch
        out['sparse_percentile'] = ppgs.sparsify(batch.clone(), 'percentile', torch.tensor([0.85]))
        out['sparse_percentile50'] = ppgs.sparsify(batch.clone(), 'percentile', torch.tensor([0.5]))
        out['sparse_constant'] = ppgs.sparsify(batch.clone(), 'constant', torch.tensor([0.1]))
        out['sparse_topk'] = ppgs.sparsify(batch.clone(), 'topk', 3)
        g = torch.Generator().manual_seed(seed)
        interp = torch.rand(frames, generator=g)
        out['interp'] = interp
        out['interpolate_vector'] = ppgs.interpolate(x, y, interp)
        out['interpolate_scalar'] = ppgs.interpolate(x, y, 0.3)
        grid = torch.cat((torch.rand(frames * 2, generator=g) * (frames - 1),
                          torch.tensor([0.0, frames - 1.0, frames - 1.5])))
        out['grid'] = grid
        out['grid_sample'] = ppgs.edit.grid.sample(x, grid)
        path = os.path.join(ROOT, 'tests', 'golden', f'postops_s{seed}.npz')
        np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
        print(path, {k: tuple(v.shape) for k, v in out.items()})


if __name__ == '__main__':
    main()
