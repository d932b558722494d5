"""Cycle accounting of the GEMM pipeline roles (PPGS_B200_TRACE=1): who waits for whom."""
import ctypes
import os
import sys

os.environ['PPGS_B200_TRACE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import ppgs_b200  # noqa: E402
from ppgs_b200 import _lib  # noqa: E402
from oracle import ppg_oracle as O  # noqa: E402

engine = ppgs_b200.Engine(0).load_state_dict(O.random_state_dict(0, peaky=True))
engine.precision = 'f16x2'
audio = O.synthetic_audio(64, 160000, 0).cuda()
engine.from_audio(audio)
buf = (ctypes.c_ulonglong * 128)()
_lib.lib.ppgs_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
_lib.check(_lib.lib.ppgs_debug_trace(engine._handle, buf))
forwards = 3
for _ in range(forwards):
    engine.from_audio(audio)
_lib.check(_lib.lib.ppgs_debug_trace(engine._handle, buf))
names = ['conv_in', 'qkv', 'out_proj', 'ffn1', 'ffn2']
launches = [1, 5, 5, 5, 5]
print('per launch, averaged over CTAs, in kcycles: mma[wait_tmem_empty wait_full total] '
      'epi[wait_tmem_full total] prod[wait_empty total]')
for k, name in enumerate(names):
    c = [buf[8 * k + i] for i in range(8)]
    ctas = max(c[7], 1)
    mma_ctas = ctas / 2 if engine else ctas   # pair mode: one MMA issuer per pair
    f = lambda x, n: x / n / 1e3   # noqa: E731
    print(f'{name:9s} mma[{f(c[0], mma_ctas):8.1f} {f(c[1], mma_ctas):8.1f} {f(c[2], mma_ctas):8.1f}] '
          f'epi[{f(c[3], ctas):8.1f} {f(c[4], ctas):8.1f}] prod[{f(c[5], ctas):8.1f} {f(c[6], ctas):8.1f}] '
          f'ctas/launch {ctas / forwards / launches[k]:.0f}')

print('LayerNorm epilogue phases, cycles per tile (warp 2 lane 0): pass1 exch1 pass2 exch2 pass3')
for slot, name in ((6, 'out_proj'), (7, 'ffn2')):
    c = [buf[8 * slot + i] for i in range(8)]
    n = max(c[5], 1)
    print(f'{name:9s} ' + ' '.join(f'{c[i] / n:8.0f}' for i in range(5)) + f'   tiles {n}')

c = [buf[64 + i] for i in range(16)]
if c[13]:
    n_cta, n_mma = c[13], c[13] / 2
    print('fused FFN per launch per CTA (kcycles): MMA waits x_full hacc_empty w1_full h1_full w2_full '
          'y_empty | total')
    print('   ' + ' '.join(f'{c[i] / n_mma / 1e3:8.1f}' for i in range(7)))
    print('epilogue waits hacc_full h1_empty y_full | LayerNorm | total')
    print('   ' + ' '.join(f'{c[8 + i] / n_cta / 1e3:8.1f}' for i in range(5)))

c = [buf[80 + i] for i in range(16)]
if c[14] and os.environ.get('PPGS_B200_ATTN_DUAL', '1') != '0':
    n = c[14]
    print(f'two-tile attention per CTA (cycles, {n} CTAs), lane 0: MMA waits q | s_empty | k_full | p_full | v_full | total')
    print('   ' + ' '.join(f'{c[i] / n:8.0f}' for i in range(6)))
    print('softmax warp: wait s_full | wait o_done | block loop | epilogue | kernel start -> first scores | kernel total')
    print('   ' + ' '.join(f'{c[8 + i] / n:8.0f}' for i in range(6)))
elif c[14]:
    n = c[14]
    print(f'attention per CTA (cycles, {n} CTAs): MMA waits q k p v | total')
    print('   ' + ' '.join(f'{c[i] / n:8.0f}' for i in range(5)))
    print('softmax warp: wait S | row max | wait p_empty | chunk work | wait O | total')
    print('   ' + ' '.join(f'{c[8 + i] / n:8.0f}' for i in range(6)))
