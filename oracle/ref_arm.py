"""The reference's own inference path, run as shipped, for the benchmark's baseline arms.

BENCH INFRASTRUCTURE ONLY (never imported by the product).  Imports the UNMODIFIED reference
package (`oracle/_ref/ppgs`, the travelled copy made by `oracle/build_ref.py`, or
`/root/reference/ppgs` in the dev container) under `oracle/refshim.py` and drives the call
pattern of its batched inference loop, `ppgs/core.py:333-352` (`from_dataloader`):

    features = ppgs.preprocess.mel.from_audios(audio, lengths, gpu=gpu)      # preprocess/mel.py:14-19
    ppgs.from_features(features, frame_lengths, representation='mel',
                       checkpoint=ckpt, gpu=gpu)                             # core.py:72-128 -> infer :551-596

`ppgs.from_audio` itself only works for batch 1, which is why the batched
configs go through `from_audios` + `from_features`, exactly like the reference's own loop.
`gpu=None` is the CPU arm (bf16 autocast via torchutil.inference.context, restated in refshim);
`gpu=0` is the same-box PyTorch-eager arm under fp16 autocast (cuBLAS / cuDNN / SDPA kernels).
"""
import os
import tempfile

import torch

from . import refshim

_STATE = {}


def available():
    return refshim.available()


def kind():
    """`reference` when the reference's own code runs, else `port` (oracle.AsShipped)."""
    return 'reference' if available() else 'port'


def _reference():
    if 'ppgs' not in _STATE:
        _STATE['ppgs'] = refshim.import_reference()
    return _STATE['ppgs']


def checkpoint_for(state_dict):
    """Write `{'model': state_dict}` like the reference's trainer does and return the path."""
    import hashlib
    digest = hashlib.sha1()
    for name in sorted(state_dict):
        digest.update(name.encode())
        digest.update(state_dict[name].detach().cpu().contiguous().numpy().tobytes())
    key = digest.hexdigest()   # by content: `id()` of a freed dict can be reused by the next one
    if key not in _STATE:
        tmp = tempfile.NamedTemporaryFile(prefix='ppgs_ref_', suffix='.pt', delete=False)
        tmp.close()
        torch.save({'model': {k: v.clone() for k, v in state_dict.items()}}, tmp.name)
        _STATE[key] = tmp.name
    return _STATE[key]


def from_audios(audio, checkpoint, gpu=None, lengths=None):
    """audio (B,1,samples) fp32 CPU tensor -> posteriors (B,40,frames) as the reference
    returns them (bf16 on CPU, fp32 softmax of fp16 logits on CUDA), on the reference's device."""
    ppgs = _reference()
    if lengths is None:
        lengths = torch.full((audio.shape[0],), audio.shape[-1], dtype=torch.long)
    features = ppgs.preprocess.mel.from_audios(audio, lengths, gpu=gpu)
    frame_lengths = (lengths // ppgs.HOPSIZE).to(torch.long)
    return ppgs.from_features(features, frame_lengths, representation='mel',
                              checkpoint=checkpoint, gpu=gpu)


def modules_fp32(checkpoint, device):
    """The reference's own `ppgs.Model` in fp32 with autocast off (oracle mode O3) on `device`:
    returns f(audio (B,1,samples) on any device) -> posteriors fp32."""
    ppgs = _reference()
    model = ppgs.load.model(checkpoint=checkpoint, representation='mel').to(device).eval()
    gpu = None if torch.device(device).type == 'cpu' else torch.device(device).index or 0

    def run(audio, lengths=None):
        if lengths is None:
            lengths = torch.full((audio.shape[0],), audio.shape[-1], dtype=torch.long)
        with torch.inference_mode():
            features = ppgs.preprocess.mel.from_audios(audio, lengths, gpu=gpu)
            frames = (lengths // ppgs.HOPSIZE).to(device)
            return torch.softmax(model(features.float(), frames), dim=1)
    return run
