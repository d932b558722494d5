"""The post-processing oracle (oracle/postops_oracle.py) against golden outputs of the
reference's own ppgs.distance / sparsify / interpolate / edit.grid.sample
(oracle/make_golden_postops.py), and live against the reference when its tree exists."""
import os

import pytest
import torch

from conftest import golden
from oracle import postops_oracle as P

CASES = ['postops_s0', 'postops_s1']


def tensors(name):
    return {k: torch.from_numpy(v) for k, v in golden(name).items()}


@pytest.mark.parametrize('name', CASES)
def test_postops_oracle_vs_reference_golden(name):
    g = tensors(name)
    x, y, similarity = g['x'], g['y'], g['similarity']
    for reduction in ('mean', 'sum', 'none'):
        got = P.distance(x, y, reduction=reduction, similarity=similarity)
        assert torch.allclose(got, g[f'distance_{reduction}'], rtol=1e-5, atol=1e-6)
        got = P.distance(x, y, reduction=reduction, normalize=False)
        assert torch.allclose(got, g[f'distance_raw_{reduction}'], rtol=1e-5, atol=1e-6)
    assert torch.allclose(P.distance(x, y, exponent=2.0, similarity=similarity), g['distance_exp2'],
                          rtol=1e-5)
    batch = g['batch']
    # the reference's 1-element-tensor threshold adds a leading axis (torch.quantile)
    assert torch.equal(P.sparsify(batch, 'percentile', 0.85), g['sparse_percentile'][0])
    assert torch.equal(P.sparsify(batch, 'percentile', 0.5), g['sparse_percentile50'][0])
    assert torch.equal(P.sparsify(batch, 'constant', 0.1), g['sparse_constant'])
    assert torch.equal(P.sparsify(batch, 'topk', 3), g['sparse_topk'])
    assert torch.equal(P.interpolate(x, y, g['interp']), g['interpolate_vector'])
    assert torch.equal(P.interpolate(x, y, 0.3), g['interpolate_scalar'])
    assert torch.equal(P.grid_sample(x, g['grid']), g['grid_sample'])


@pytest.mark.skipif(not os.path.isdir('/root/reference/ppgs'),
                    reason='reference tree only exists in the dev container')
def test_postops_oracle_vs_live_reference():
    from oracle import refshim
    ppgs = refshim.import_reference()
    x, y = P.random_ppg(21, 77), P.random_ppg(22, 77, sharpness=0.5)
    similarity = P.random_similarity(9)
    ppgs.distance.similarity_matrix, ppgs.distance.device = similarity, x.device
    assert torch.allclose(ppgs.distance(x, y), P.distance(x, y, similarity=similarity), rtol=1e-5)
    batch = P.random_ppg(23, 50, batch=1)
    assert torch.equal(ppgs.sparsify(batch.clone(), 'topk', 5), P.sparsify(batch, 'topk', 5))
    grid = torch.linspace(0, 76, 200)
    assert torch.equal(ppgs.edit.grid.sample(x, grid), P.grid_sample(x, grid))
