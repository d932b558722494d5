"""Grid-based PPG interpolation (ppgs/edit/grid.py:13-85) — the gather + linear
interpolation runs as one CUDA kernel (ppgs_ppg_grid_sample)."""
import ctypes

import torch

from .. import _lib
from .. import load


def sample(ppg: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """Grid-based PPG interpolation (ppgs/edit/grid.py:13-50).

    ppg: (..., frames); grid: float-valued frame indices, shape (samples,);
    returns (..., samples) where `ppg` lives."""
    engine = load.utility_engine(ppg.device.index if ppg.is_cuda else None)
    x = ppg.to(engine.device, torch.float32).contiguous()
    g = grid.to(engine.device, torch.float32).contiguous()
    if g.dim() != 1:
        raise ValueError('grid must have shape (samples,)')
    frames = x.shape[-1]
    rows = x.numel() // max(frames, 1)
    out = torch.empty(x.shape[:-1] + (g.shape[0],), dtype=torch.float32, device=engine.device)
    _lib.check(_lib.lib.ppgs_ppg_grid_sample(
        engine._handle, ctypes.c_void_p(x.data_ptr()), rows, frames, ctypes.c_void_p(g.data_ptr()),
        g.shape[0], ctypes.c_void_p(out.data_ptr()),
        ctypes.c_void_p(torch.cuda.current_stream(engine.device).cuda_stream)))
    out = out.to(ppg.dtype)
    return out if ppg.is_cuda else out.cpu()


def constant(ppg: torch.Tensor, ratio: float) -> torch.Tensor:
    """Grid for constant-ratio time-stretching (ppgs/edit/grid.py:53-66); lower is slower."""
    return of_length(ppg, round(ppg.shape[-1] / ratio + 1e-4))


def of_length(ppg: torch.Tensor, length: int) -> torch.Tensor:
    """Grid that resamples a PPG to `length` frames (ppgs/edit/grid.py:108-125)."""
    return torch.linspace(
        0., ppg.shape[-1] - 1., length, dtype=torch.float, device=ppg.device)
