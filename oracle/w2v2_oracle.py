"""CPU oracle of the `w2v2fb` representation: a functional restatement of
`ppgs.preprocess.w2v2fb.from_audios` (ppgs/preprocess/w2v2fb/core.py:32-75) and of the
third-party model it calls, Hugging Face `transformers.Wav2Vec2Model` with the
`facebook/wav2vec2-base` architecture (transformers is NOT vendored in the reference and is
un-pinned in its setup.py:39; restated from transformers 5.5.0
`models/wav2vec2/modeling_wav2vec2.py`: feature encoder :254-324,:382-419, feature
projection :422-436, positional conv :326-380, attention :440-550, encoder layer :576-610,
encoder :658-728, mask reduction :1005-1044, model forward :1327-1389).

TEST INFRASTRUCTURE ONLY (same rules as oracle/ppg_oracle.py).  Pinned against the real
`transformers.Wav2Vec2Model` run in the dev container on seeded random weights
(tests/test_w2v2_oracle.py and the golden fixture tests/golden/w2v2fb_*.npz made by
oracle/make_golden_w2v2.py).  Pretrained weights are not available offline.
"""
import math

import torch

HOPSIZE = 160
W2V2_PAD = 40            # WINDOW_SIZE // 2 - HOP_SIZE // 2 = 200 - 160 (w2v2fb/core.py:17-22,54)
CONV_DIM = 512
CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)
HIDDEN = 768
HEADS = 12
LAYERS = 12
FFN = 3072
POS_KERNEL = 128
POS_GROUPS = 16
EPS = 1e-5


def random_state_dict(seed=0):
    """Seeded weights in the Hugging Face `Wav2Vec2Model` state-dict schema (the keys
    `from_pretrained('facebook/wav2vec2-base')` produces, minus `masked_spec_embed`)."""
    g = torch.Generator().manual_seed(20_000 + seed)

    def uniform(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    sd = {}
    c_in = 1
    for i, k in enumerate(CONV_KERNEL):
        sd[f'feature_extractor.conv_layers.{i}.conv.weight'] = uniform(
            (CONV_DIM, c_in, k), math.sqrt(3.0 / (c_in * k)) * 1.4)
        c_in = CONV_DIM
    sd['feature_extractor.conv_layers.0.layer_norm.weight'] = 1 + uniform((CONV_DIM,), 0.1)
    sd['feature_extractor.conv_layers.0.layer_norm.bias'] = uniform((CONV_DIM,), 0.1)
    sd['feature_projection.layer_norm.weight'] = 1 + uniform((CONV_DIM,), 0.1)
    sd['feature_projection.layer_norm.bias'] = uniform((CONV_DIM,), 0.1)
    sd['feature_projection.projection.weight'] = uniform((HIDDEN, CONV_DIM), 1 / math.sqrt(CONV_DIM))
    sd['feature_projection.projection.bias'] = uniform((HIDDEN,), 0.05)
    sd['encoder.pos_conv_embed.conv.bias'] = uniform((HIDDEN,), 0.05)
    sd['encoder.pos_conv_embed.conv.parametrizations.weight.original0'] = 0.5 + torch.rand(
        (1, 1, POS_KERNEL), generator=g)
    sd['encoder.pos_conv_embed.conv.parametrizations.weight.original1'] = uniform(
        (HIDDEN, HIDDEN // POS_GROUPS, POS_KERNEL), 0.05)
    sd['encoder.layer_norm.weight'] = 1 + uniform((HIDDEN,), 0.1)
    sd['encoder.layer_norm.bias'] = uniform((HIDDEN,), 0.1)
    for i in range(LAYERS):
        p = f'encoder.layers.{i}.'
        for name in ('q_proj', 'k_proj', 'v_proj', 'out_proj'):
            sd[p + f'attention.{name}.weight'] = uniform((HIDDEN, HIDDEN), 1 / math.sqrt(HIDDEN))
            sd[p + f'attention.{name}.bias'] = uniform((HIDDEN,), 0.02)
        sd[p + 'layer_norm.weight'] = 1 + uniform((HIDDEN,), 0.1)
        sd[p + 'layer_norm.bias'] = uniform((HIDDEN,), 0.1)
        sd[p + 'feed_forward.intermediate_dense.weight'] = uniform((FFN, HIDDEN), 1 / math.sqrt(HIDDEN))
        sd[p + 'feed_forward.intermediate_dense.bias'] = uniform((FFN,), 0.02)
        sd[p + 'feed_forward.output_dense.weight'] = uniform((HIDDEN, FFN), 1 / math.sqrt(FFN))
        sd[p + 'feed_forward.output_dense.bias'] = uniform((HIDDEN,), 0.02)
        sd[p + 'final_layer_norm.weight'] = 1 + uniform((HIDDEN,), 0.1)
        sd[p + 'final_layer_norm.bias'] = uniform((HIDDEN,), 0.1)
    return sd


def conv_out_length(length, kernel, stride):
    return (length - kernel) // stride + 1


def feature_lengths(sample_lengths):
    """modeling_wav2vec2.py:1005-1024 applied to every conv layer."""
    out = sample_lengths.clone()
    for k, s in zip(CONV_KERNEL, CONV_STRIDE):
        out = torch.div(out - k, s, rounding_mode='floor') + 1
    return out


def gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, weight, bias):
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + EPS) * weight + bias


def conv1d_strided(x, weight, stride):
    """x (B,Cin,T), weight (Cout,Cin,k), no bias / padding -> (B,Cout,(T-k)//stride+1),
    as an explicit sum over taps of strided matmuls."""
    k = weight.shape[-1]
    t_out = (x.shape[-1] - k) // stride + 1
    out = None
    for tap in range(k):
        sl = x[..., tap:tap + stride * (t_out - 1) + 1:stride]
        term = torch.einsum('oc,bct->bot', weight[:, :, tap], sl)
        out = term if out is None else out + term
    return out


def feature_encoder(sd, audio):
    """Wav2Vec2FeatureEncoder (feat_extract_norm='group', conv_bias=False, GELU):
    layer 0 = conv + GroupNorm(512 groups == per-channel statistics over time) + GELU,
    layers 1-6 = conv + GELU.  audio (B,T) -> (B,512,T6)."""
    h = audio[:, None]
    for i, stride in enumerate(CONV_STRIDE):
        h = conv1d_strided(h, sd[f'feature_extractor.conv_layers.{i}.conv.weight'], stride)
        if i == 0:
            mean = h.mean(-1, keepdim=True)
            var = ((h - mean) ** 2).mean(-1, keepdim=True)
            h = (h - mean) / torch.sqrt(var + EPS)
            h = h * sd['feature_extractor.conv_layers.0.layer_norm.weight'][None, :, None] \
                + sd['feature_extractor.conv_layers.0.layer_norm.bias'][None, :, None]
        h = gelu(h)
    return h


def positional_conv(sd, h):
    """Wav2Vec2PositionalConvEmbedding: weight-normalised (dim=2) grouped Conv1d(768,768,
    k=128, pad 64, groups 16), drop the last frame, GELU.  h (B,T,768) -> (B,T,768)."""
    g = sd['encoder.pos_conv_embed.conv.parametrizations.weight.original0']
    v = sd['encoder.pos_conv_embed.conv.parametrizations.weight.original1']
    norm = torch.sqrt((v * v).sum(dim=(0, 1), keepdim=True))
    weight = v * (g / norm)
    x = torch.nn.functional.pad(h.transpose(1, 2), (POS_KERNEL // 2, POS_KERNEL // 2))
    B, _, T = x.shape
    t_out = T - POS_KERNEL + 1
    per = HIDDEN // POS_GROUPS
    out = torch.zeros(B, HIDDEN, t_out, dtype=h.dtype)
    for grp in range(POS_GROUPS):
        xs = x[:, grp * per:(grp + 1) * per]
        ws = weight[grp * per:(grp + 1) * per]
        acc = torch.zeros(B, per, t_out, dtype=h.dtype)
        for tap in range(POS_KERNEL):
            acc = acc + torch.einsum('oc,bct->bot', ws[:, :, tap], xs[..., tap:tap + t_out])
        out[:, grp * per:(grp + 1) * per] = acc
    out = out + sd['encoder.pos_conv_embed.conv.bias'][None, :, None]
    out = out[..., :-1]                     # Wav2Vec2SamePadLayer (even kernel)
    return gelu(out).transpose(1, 2)


def encoder_layer(sd, prefix, h, key_mask):
    """Wav2Vec2EncoderLayer (post-LN): h = LN(h + Attn(h)); h = LN(h + FF(h))."""
    B, T, H = h.shape
    d = H // HEADS

    def proj(name):
        return h @ sd[prefix + f'attention.{name}.weight'].T + sd[prefix + f'attention.{name}.bias']

    q = proj('q_proj').reshape(B, T, HEADS, d).transpose(1, 2)
    k = proj('k_proj').reshape(B, T, HEADS, d).transpose(1, 2)
    v = proj('v_proj').reshape(B, T, HEADS, d).transpose(1, 2)
    scores = (q @ k.transpose(-1, -2)) * (d ** -0.5)
    neg = torch.full((), float('-inf'), dtype=h.dtype)
    scores = torch.where(key_mask[:, None, None, :], scores, neg)
    p = torch.softmax(scores, dim=-1)
    attn = (p @ v).transpose(1, 2).reshape(B, T, H)
    attn = attn @ sd[prefix + 'attention.out_proj.weight'].T + sd[prefix + 'attention.out_proj.bias']
    h = layer_norm(h + attn, sd[prefix + 'layer_norm.weight'], sd[prefix + 'layer_norm.bias'])
    ff = gelu(h @ sd[prefix + 'feed_forward.intermediate_dense.weight'].T
              + sd[prefix + 'feed_forward.intermediate_dense.bias'])
    ff = ff @ sd[prefix + 'feed_forward.output_dense.weight'].T + sd[prefix + 'feed_forward.output_dense.bias']
    return layer_norm(h + ff, sd[prefix + 'final_layer_norm.weight'], sd[prefix + 'final_layer_norm.bias'])


def wav2vec2_forward(sd, padded_audio, valid_samples, dtype=torch.float32):
    """`Wav2Vec2Model(padded_audio, attention_mask).last_hidden_state` in eval mode.
    padded_audio (B,T); valid_samples (B,) = number of unmasked samples per row."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    feats = feature_encoder(sd, padded_audio.to(dtype)).transpose(1, 2)       # (B,T6,512)
    T6 = feats.shape[1]
    out_len = feature_lengths(valid_samples)
    key_mask = torch.arange(T6)[None] < out_len[:, None]                      # :1026-1044
    h = layer_norm(feats, sd['feature_projection.layer_norm.weight'],
                   sd['feature_projection.layer_norm.bias'])
    h = h @ sd['feature_projection.projection.weight'].T + sd['feature_projection.projection.bias']
    h = h * key_mask[..., None].to(dtype)                                     # :681-684
    h = h + positional_conv(sd, h)
    h = layer_norm(h, sd['encoder.layer_norm.weight'], sd['encoder.layer_norm.bias'])
    for i in range(LAYERS):
        h = encoder_layer(sd, f'encoder.layers.{i}.', h, key_mask)
    return h


def nearest_upsample(h, frames):
    """F.interpolate(mode='nearest', size=frames) over time: src = floor(dst * (in / out))
    with the scale computed in fp32, as ATen does.  h (B,C,T) -> (B,C,frames)."""
    T = h.shape[-1]
    scale = torch.tensor(T, dtype=torch.float32) / torch.tensor(frames, dtype=torch.float32)
    index = torch.floor(torch.arange(frames, dtype=torch.float32) * scale).long().clamp(max=T - 1)
    return h[..., index]


def from_audios(sd, audio, lengths, dtype=torch.float32):
    """ppgs/preprocess/w2v2fb/core.py:32-75 at 16 kHz.  audio (B,1,samples) zero-padded,
    lengths (B,) samples -> (B,768,samples//160) fp16."""
    padded = torch.nn.functional.pad(audio, (W2V2_PAD, W2V2_PAD)).squeeze(1)
    hidden = wav2vec2_forward(sd, padded, lengths + 2 * W2V2_PAD, dtype)
    up = nearest_upsample(hidden.transpose(1, 2), audio.shape[-1] // HOPSIZE)
    return up.to(torch.float16)
