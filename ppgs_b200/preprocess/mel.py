"""mel representation (ppgs/preprocess/mel.py:14-30, spectrogram.py:14-50) on
the fused CUDA kernel.  Numerics: fp32 STFT / filterbank with the reference's two
fp16 roundings — the autocast-off variant the checkpoints were trained on
(SURVEY.md F6, F7)."""
import torch

from .. import config
from .. import load


def _frontend_engine(gpu):
    # any finalised engine on the device owns the mel tables; prefer a cached one
    with load._lock:
        for key, engine in load._engines.items():
            if key[0] != 'utility' and key[2] == load.resolve_device(gpu).index:
                return engine
    return _standalone(load.resolve_device(gpu))


_standalone_engines = {}


def _standalone(device):
    """Front-end-only engine (tiny 1-layer model with zero weights) for callers
    that want mels without loading a checkpoint."""
    engine = _standalone_engines.get(device.index)
    if engine is None:
        from ..engine import Engine
        engine = Engine(device, num_hidden_layers=1)
        cfg = engine.cfg
        H, C, F, O, k = (cfg.hidden_channels, cfg.input_channels, cfg.ffn_channels,
                         cfg.output_channels, cfg.kernel_size)
        z = torch.zeros
        state = {
            'position.encoding': z(cfg.max_len, 1, H),
            'input_layer.weight': z(H, C, k), 'input_layer.bias': z(H),
            'output_layer.weight': z(O, H, k), 'output_layer.bias': z(O),
        }
        p = 'model.layers.0.'
        state.update({
            p + 'self_attn.in_proj_weight': z(3 * H, H), p + 'self_attn.in_proj_bias': z(3 * H),
            p + 'self_attn.out_proj.weight': z(H, H), p + 'self_attn.out_proj.bias': z(H),
            p + 'linear1.weight': z(F, H), p + 'linear1.bias': z(F),
            p + 'linear2.weight': z(H, F), p + 'linear2.bias': z(H),
            p + 'norm1.weight': z(H), p + 'norm1.bias': z(H),
            p + 'norm2.weight': z(H), p + 'norm2.bias': z(H)})
        engine.load_state_dict(state)
        _standalone_engines[device.index] = engine
    return engine


def from_audios(audio, lengths=None, sample_rate=config.SAMPLE_RATE, gpu=None):
    """(B,1,samples) fp32 -> (B,80,samples//160) fp16 on the GPU."""
    return _frontend_engine(gpu).mel(audio)


def from_audio(audio, sample_rate=config.SAMPLE_RATE, gpu=None):
    if audio.dim() == 2:
        audio = audio.unsqueeze(dim=0)
    return from_audios(audio, lengths=audio.shape[-1], sample_rate=sample_rate, gpu=gpu)
