"""Single-kernel checks of the tensor-core building blocks through the library's
validation entry points (ppgs_b200/csrc/debug_abi.h): the tcgen05 GEMM (TMA
128B-swizzled operand ring, split-fp16 passes, TMEM accumulator) and the
attention kernel, against fp64 torch on the same fp32 inputs."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import ppg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engine():
    import ppgs_b200
    return ppgs_b200.Engine(0).load_state_dict(O.random_state_dict(0))


def debug_lib():
    from ppgs_b200 import _lib
    lib = _lib.lib
    vp, i = ctypes.c_void_p, ctypes.c_int
    lib.ppgs_debug_gemm.restype = i
    lib.ppgs_debug_gemm.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, i, i, vp]
    lib.ppgs_debug_attention.restype = i
    lib.ppgs_debug_attention.argtypes = [vp, vp, i, i, i, i, i, i, i, i, vp]
    return _lib


def run_gemm(engine, a, w, bias, taps, bn, a_planes, b_planes, pair=0):
    lib = debug_lib()
    M, C = a.shape
    N = w.shape[0]
    out = torch.empty(M, N)
    a, w, bias = a.contiguous(), w.contiguous(), bias.contiguous()
    lib.check(lib.lib.ppgs_debug_gemm(
        engine._handle, a.data_ptr(), w.data_ptr(), bias.data_ptr(), M, N, C, taps, bn, pair,
        a_planes, b_planes, out.data_ptr()))
    return out


def reference_gemm(a, w, bias, taps):
    a64, w64 = a.double(), w.double().reshape(w.shape[0], a.shape[1], taps)
    half = taps // 2
    padded = torch.nn.functional.pad(a64, (0, 0, half, half))
    out = bias.double()[None].repeat(a.shape[0], 1)
    for tap in range(taps):
        out += padded[tap:tap + a.shape[0]] @ w64[:, :, tap].T
    return out


GEMM_CASES = [
    # M, N, C, taps, bn[, pair]
    (256, 256, 64, 1, 256, 1),
    (256, 768, 256, 1, 256, 1),
    (512, 2048, 256, 1, 256, 1),
    (256, 256, 2048, 1, 256, 1),
    (256, 256, 80, 5, 256, 1),
    (256 * 160, 256, 128, 1, 256, 1),  # CTA pairs: more pair-tiles than clusters
    (128, 256, 64, 1, 256),
    (128, 256, 256, 1, 256),
    (384, 768, 256, 1, 256),
    (256, 2048, 256, 1, 256),
    (256, 256, 2048, 1, 256),
    (256, 128, 192, 1, 128),
    (256, 40, 256, 5, 64),
    (256, 256, 80, 5, 256),
    (128 * 310, 256, 64, 1, 256),     # more tiles than SMs: persistent loop, TMEM phases
]


@pytest.mark.parametrize('case', GEMM_CASES, ids=lambda c: 'x'.join(map(str, c)))
@pytest.mark.parametrize('planes', [(2, 2), (1, 2), (1, 1)])
def test_tcgen05_gemm(engine, case, planes):
    M, N, C, taps, bn = case[:5]
    pair = case[5] if len(case) > 5 else 0
    g = torch.Generator().manual_seed(M + N + C + taps)
    a = torch.randn(M, C, generator=g)
    if planes[0] == 1:
        a = a.half().float()          # exact fp16 A (the input conv's case)
    w = torch.randn(N, C, taps, generator=g) / (C * taps) ** 0.5
    bias = torch.randn(N, generator=g)
    out = run_gemm(engine, a, w, bias, taps, bn, *planes, pair=pair)
    ref = reference_gemm(a, w, bias, taps)
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    # split-fp16 passes keep ~22 bits per operand; the fp32 TMEM accumulation adds
    # O(K) rounding.  Single-pass fp16 operands: ~2^-11 per product.
    tol = 1e-5 if planes[1] == 2 else 2e-3
    assert err <= tol, f'relative error {err:.3e}'


def attention_reference(qkv, rows, H, heads, valid_len, causal):
    q, k, v = [t.double().reshape(rows, heads, H // heads).transpose(0, 1) for t in qkv.split(H, 1)]
    scores = q @ k.transpose(1, 2) / (H // heads) ** 0.5
    scores[:, :, valid_len:] = float('-inf')
    if causal:
        scores = scores.masked_fill(torch.ones(rows, rows).triu(1).bool(), float('-inf'))
    p = torch.softmax(scores, -1) if valid_len else torch.zeros_like(scores)
    return (p @ v).transpose(0, 1).reshape(rows, H)


@pytest.mark.parametrize('tensor_len,valid_len', [(126, 126), (500, 500), (500, 317), (250, 1), (300, 0),
                                                  (380, 300), (1200, 1111)])
@pytest.mark.parametrize('causal', [0, 1])
@pytest.mark.parametrize('growth', [0.0, 3.0])
def test_attention_two_tile_kernel(engine, tensor_len, valid_len, causal, growth):
    """The default model's attention (head_dim 128, attention_dual_tc.cu): two query tiles per
    CTA over shared K / V blocks, Q / K / P as single fp16 planes, online softmax.  Odd tile
    counts (380 -> 3 tiles: one idle lane), sequences longer than the old 512-key limit, causal
    masks (the earlier tile needs fewer key blocks than its partner) and key norms that GROW
    along the sequence (`growth`), so that later blocks exceed the running max by far more than
    the 2^8 slack and the accumulator is rescaled in TMEM.  Tolerance: the fp16 rounding of
    Q, K (scores) and P, i.e. the reference's own fp16-autocast numerics for this product."""
    lib = debug_lib()
    H, heads = 256, 2
    rows = (tensor_len + 2 + 127) // 128 * 128
    g = torch.Generator().manual_seed(tensor_len + valid_len + causal)
    qkv = torch.randn(rows, 3 * H, generator=g)
    qkv[:, :H] *= 2.0
    qkv[:, H:2 * H] *= (1.0 + growth * torch.arange(rows) / rows)[:, None] ** 2
    out = torch.empty(rows, H)
    lib.check(lib.lib.ppgs_debug_attention(
        engine._handle, qkv.data_ptr(), rows, tensor_len, valid_len, H, heads, causal, 2, 2, out.data_ptr()))
    # reference on the operands the kernel is specified to see: Q, K rounded to fp16
    seen = qkv.clone()
    seen[:, :2 * H] = seen[:, :2 * H].half().float()
    ref = attention_reference(seen, rows, H, heads, valid_len, causal)
    err = (out.double() - ref)[:tensor_len].abs().max().item()
    assert err <= 2e-3, f'max-abs {err:.3e}'
    if not growth:   # and against the unrounded operands at the fp16-autocast class of error
        exact = attention_reference(qkv, rows, H, heads, valid_len, causal)
        assert (out.double() - exact)[:tensor_len].abs().max().item() <= 5e-3


@pytest.mark.parametrize('tensor_len,valid_len', [(126, 126), (500, 500), (500, 317), (250, 1), (300, 0)])
@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('H,heads,causal', [(256, 2, 0), (128, 2, 0), (512, 2, 0), (768, 12, 0), (256, 2, 1),
                                            (512, 2, 1)])
def test_attention_kernel(engine, tensor_len, valid_len, impl, H, heads, causal):
    """head_dim 128 (mel PPG model), 64 (wav2vec2 encoder) and 256 (w2v2fb PPG model)."""
    lib = debug_lib()
    rows = (tensor_len + 2 + 127) // 128 * 128
    g = torch.Generator().manual_seed(tensor_len + valid_len)
    qkv = torch.randn(rows, 3 * H, generator=g)
    qkv[:, :H] *= 2.0
    out = torch.empty(rows, H)
    rc = lib.lib.ppgs_debug_attention(
        engine._handle, qkv.data_ptr(), rows, tensor_len, valid_len, H, heads, causal, 2, impl,
        out.data_ptr())
    lib.check(rc)
    q, k, v = [t.double().reshape(rows, heads, H // heads).transpose(0, 1) for t in qkv.split(H, 1)]
    scores = q @ k.transpose(1, 2) / (H // heads) ** 0.5
    scores[:, :, valid_len:] = float('-inf')
    if causal:
        scores = scores.masked_fill(torch.ones(rows, rows).triu(1).bool(), float('-inf'))
    p = torch.softmax(scores, -1) if valid_len else torch.zeros_like(scores)
    ref = (p @ v).transpose(0, 1).reshape(rows, H)
    err = (out.double() - ref)[:tensor_len].abs().max().item()
    assert err <= 2e-5, f'max-abs {err:.3e}'
