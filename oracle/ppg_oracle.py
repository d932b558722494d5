"""CPU oracle: a functional restatement of the reference `ppgs.from_audio` path.

TEST INFRASTRUCTURE ONLY — this file is the *checker*, never the product.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import it.  `ppgs_b200` never imports anything under `oracle/`; its hot path is
the CUDA library and it raises when that library is missing.

Parity pin: the reference ships NO tests, golden vectors or fixtures
(SURVEY.md §4), so this oracle is pinned against *outputs of the reference's own
modules run in the dev container* under `oracle/refshim.py` (script:
`oracle/make_golden.py`; committed fixtures: `tests/golden/*.npz`; live check:
`tests/test_oracle_golden.py::test_oracle_vs_live_reference`, which runs whenever `/root/reference`
exists).  Third-party boundaries that remain unpinned: `torchutil.inference
.context` (restated, package absent) and `librosa.filters.mel` (restated from the
published Slaney construction, checked against torchaudio's implementation).

Each function cites the reference lines it follows.  Everything is plain torch
CPU tensor arithmetic (fp32 by default, fp64 on request) written as explicit
matmuls/loops — no `nn.Module` — except `AsShipped`, which rebuilds the
reference's module stack from stock torch layers so that the CPU baseline timed
in `bench.py` executes the same library kernels (bf16 autocast) the reference
does on `gpu=None`.
"""
import math

import numpy as np
import torch

# Constants: ppgs/config/defaults.py:20-32,127-161; ppgs/config/w2v2fb.py:7-10
HOPSIZE = 160
NUM_FFT = 1024
WINDOW_SIZE = 1024
NUM_MELS = 80
SAMPLE_RATE = 16000
ATTENTION_HEADS = 2
HIDDEN_CHANNELS = 256
INPUT_CHANNELS = 80
KERNEL_SIZE = 5
NUM_HIDDEN_LAYERS = 5
OUTPUT_CHANNELS = 40
CHUNK_OVERLAP = 50
CHUNK_LENGTH = 500
FFN_CHANNELS = 2048      # torch.nn.TransformerEncoderLayer default dim_feedforward
LAYER_NORM_EPS = 1e-5    # torch.nn.TransformerEncoderLayer default layer_norm_eps
MAX_LEN = 5000           # ppgs/model/transformer.py:24


###############################################################################
# Synthetic inputs
###############################################################################


def synthetic_audio(batch, samples, seed=0):
    """Uniform +-0.5 noise, (batch, 1, samples) fp32."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(batch, 1, samples, generator=g) * 2 - 1) * 0.5


def speechlike_audio(batch, samples, seed=0):
    """Decaying harmonic bursts + noise floor; wider dynamic range than noise."""
    g = torch.Generator().manual_seed(1000 + seed)
    t = torch.arange(samples, dtype=torch.float64) / SAMPLE_RATE
    out = torch.zeros(batch, 1, samples, dtype=torch.float64)
    for b in range(batch):
        f0 = 90 + 120 * torch.rand(1, generator=g).item()
        env = 0.5 * (1 + torch.sin(2 * math.pi * (2.0 + b) * t)) ** 2 / 4
        sig = torch.zeros_like(t)
        for h in range(1, 30):
            amp = 1.0 / h ** (0.8 + 0.4 * torch.rand(1, generator=g).item())
            sig = sig + amp * torch.sin(2 * math.pi * f0 * h * t + h)
        noise = torch.randn(samples, generator=g, dtype=torch.float64) * 1e-3
        out[b, 0] = 0.2 * env * sig + noise
    return out.clamp(-1, 1).float()


def random_state_dict(
    seed=0,
    input_channels=INPUT_CHANNELS,
    hidden_channels=HIDDEN_CHANNELS,
    num_hidden_layers=NUM_HIDDEN_LAYERS,
    output_channels=OUTPUT_CHANNELS,
    kernel_size=KERNEL_SIZE,
    peaky=False,
):
    """Seeded weights in the reference's state-dict schema.

    Scales follow the default torch initialisers (xavier-uniform in_proj,
    kaiming-uniform(a=sqrt(5)) linears/convs) so activations look like an
    untrained reference model; `peaky` multiplies the output conv by 6
    (max posterior ~0.8) to stress precision.  Drawn from an explicit CPU
    generator so the GPU box reproduces them without the reference.
    """
    g = torch.Generator().manual_seed(10_000 + seed)
    H, C, F = hidden_channels, input_channels, FFN_CHANNELS

    def uniform(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    sd = {'position.encoding': positional_encoding(H, MAX_LEN)}
    bound = 1 / math.sqrt(C * kernel_size)
    sd['input_layer.weight'] = uniform((H, C, kernel_size), bound)
    sd['input_layer.bias'] = uniform((H,), bound)
    for layer in range(num_hidden_layers):
        p = f'model.layers.{layer}.'
        sd[p + 'self_attn.in_proj_weight'] = uniform(
            (3 * H, H), math.sqrt(6 / (4 * H)))
        sd[p + 'self_attn.in_proj_bias'] = uniform((3 * H,), 0.02)
        sd[p + 'self_attn.out_proj.weight'] = uniform((H, H), 1 / math.sqrt(H))
        sd[p + 'self_attn.out_proj.bias'] = uniform((H,), 0.02)
        sd[p + 'linear1.weight'] = uniform((F, H), 1 / math.sqrt(H))
        sd[p + 'linear1.bias'] = uniform((F,), 1 / math.sqrt(H))
        sd[p + 'linear2.weight'] = uniform((H, F), 1 / math.sqrt(F))
        sd[p + 'linear2.bias'] = uniform((H,), 1 / math.sqrt(F))
        sd[p + 'norm1.weight'] = 1 + uniform((H,), 0.1)
        sd[p + 'norm1.bias'] = uniform((H,), 0.1)
        sd[p + 'norm2.weight'] = 1 + uniform((H,), 0.1)
        sd[p + 'norm2.bias'] = uniform((H,), 0.1)
    bound = 1 / math.sqrt(H * kernel_size)
    sd['output_layer.weight'] = uniform((output_channels, H, kernel_size), bound)
    sd['output_layer.bias'] = uniform((output_channels,), bound)
    if peaky:
        sd['output_layer.weight'] = sd['output_layer.weight'] * 6
    return sd


###############################################################################
# Mel front-end
###############################################################################


def hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(
        f >= min_log_hz,
        min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep,
        mels)


def mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(
        m >= min_log_mel,
        min_log_hz * np.exp(logstep * (m - min_log_mel)),
        f_sp * m)


def mel_basis(sr=SAMPLE_RATE, n_fft=NUM_FFT, n_mels=NUM_MELS):
    """Slaney mel filterbank, (n_mels, n_fft//2+1) float32.

    Restates `librosa.filters.mel(sr=16000, n_fft=1024, n_mels=80)` (htk=False,
    norm='slaney', fmin=0, fmax=sr/2) — call site ppgs/preprocess/mel.py:61-64;
    librosa is absent here and un-pinned in the reference's setup.py:30.
    """
    fftfreqs = np.linspace(0, sr / 2, n_fft // 2 + 1)
    mel_f = mel_to_hz_slaney(
        np.linspace(hz_to_mel_slaney(0.0), hz_to_mel_slaney(sr / 2), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    # librosa keeps `weights` in float32: the triangles are rounded to fp32 on
    # assignment and the in-place Slaney scaling rounds once more.
    weights = np.zeros((n_mels, n_fft // 2 + 1), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights


def spectrogram_from_audios(audio, dtype=torch.float32):
    """ppgs/preprocess/spectrogram.py:14-50.

    audio (B,1,samples) -> magnitude spectrogram (B,513,samples//160) fp16.
    Reflect-pad (1024-160)//2 = 432, STFT n_fft=1024 hop=160 periodic hann,
    center=False, sqrt(re^2+im^2+1e-6), cast to fp16.
    """
    audio = audio.to(dtype)
    window = torch.hann_window(WINDOW_SIZE, dtype=dtype)
    size = (NUM_FFT - HOPSIZE) // 2
    audio = torch.nn.functional.pad(audio, (size, size), mode='reflect')
    stft = torch.stft(
        audio.squeeze(1), NUM_FFT, hop_length=HOPSIZE, window=window,
        center=False, normalized=False, onesided=True, return_complex=True)
    stft = torch.view_as_real(stft)
    spectrogram = torch.sqrt(stft.pow(2).sum(-1) + 1e-6)
    return spectrogram.to(torch.float16)


def mel_from_audios(audio, dtype=torch.float32):
    """ppgs/preprocess/mel.py:14-19,56-76 with autocast OFF (the dataloader /
    training numerics): fp32 filterbank matmul on the fp16-rounded
    magnitudes, log(clamp(., 1e-5)), cast back to fp16 (twice, both no-ops after
    the first).  audio (B,1,samples) -> (B,80,samples//160) fp16.
    """
    spec = spectrogram_from_audios(audio, dtype)
    basis = torch.from_numpy(mel_basis())
    mel = torch.matmul(basis.to(dtype), spec.to(dtype))
    return torch.log(torch.clamp(mel, min=1e-5)).to(torch.float16)


###############################################################################
# Transformer
###############################################################################


def positional_encoding(channels, max_len=MAX_LEN):
    """ppgs/model/transformer.py:89-101 -> (max_len, 1, channels) fp32."""
    index = torch.arange(max_len).unsqueeze(1)
    frequency = torch.exp(
        torch.arange(0, channels, 2) * (-math.log(10000.0) / channels))
    encoding = torch.zeros(max_len, 1, channels)
    encoding[:, 0, 0::2] = torch.sin(index * frequency)
    encoding[:, 0, 1::2] = torch.cos(index * frequency)
    return encoding


def mask_from_lengths(lengths):
    """ppgs/model/transformer.py:108-114 (padding=0) -> (B, max(lengths)) bool."""
    x = torch.arange(int(lengths.max()), dtype=lengths.dtype)
    return x.unsqueeze(0) < lengths.unsqueeze(1)


def conv1d_same(x, weight, bias):
    """Conv1d(k odd, padding='same') as an explicit sum of shifted matmuls.
    x (B,Cin,T), weight (Cout,Cin,k) -> (B,Cout,T)."""
    k = weight.shape[-1]
    half = k // 2
    T = x.shape[-1]
    padded = torch.nn.functional.pad(x, (half, half))
    out = bias[None, :, None].expand(x.shape[0], -1, T).clone()
    for tap in range(k):
        out = out + torch.einsum('oc,bct->bot', weight[:, :, tap], padded[..., tap:tap + T])
    return out


def layer_norm(x, weight, bias, eps=LAYER_NORM_EPS):
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * weight + bias


def encoder_layer(x, sd, prefix, heads, key_mask, causal):
    """One post-LN `torch.nn.TransformerEncoderLayer` (relu, eval):
    x = norm1(x + SA(x)); x = norm2(x + W2 relu(W1 x)).
    x (B,T,H); key_mask (B,T) bool True=valid key.
    (torch/nn/modules/transformer.py:952-956; MultiheadAttention in-proj /
    scaled dot product / out-proj.)
    """
    B, T, H = x.shape
    d = H // heads
    w_in = sd[prefix + 'self_attn.in_proj_weight']
    b_in = sd[prefix + 'self_attn.in_proj_bias']
    qkv = x @ w_in.T + b_in
    q, k, v = qkv.split(H, dim=-1)
    q = q.reshape(B, T, heads, d).transpose(1, 2)
    k = k.reshape(B, T, heads, d).transpose(1, 2)
    v = v.reshape(B, T, heads, d).transpose(1, 2)
    scores = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    neg = torch.full((), float('-inf'), dtype=x.dtype)
    allowed = key_mask[:, None, None, :].expand(B, heads, T, T)
    if causal:
        tri = torch.ones(T, T, dtype=torch.bool).tril()
        allowed = allowed & tri[None, None]
    scores = torch.where(allowed, scores, neg)
    row_max = scores.max(-1, keepdim=True).values
    row_max = torch.where(torch.isinf(row_max), torch.zeros_like(row_max), row_max)
    p = torch.exp(scores - row_max)
    denom = p.sum(-1, keepdim=True)
    # fully masked query rows -> zeros (torch's masked softmax yields 0 here)
    p = torch.where(denom > 0, p / denom.clamp(min=1e-38), torch.zeros_like(p))
    attn = (p @ v).transpose(1, 2).reshape(B, T, H)
    sa = attn @ sd[prefix + 'self_attn.out_proj.weight'].T + sd[prefix + 'self_attn.out_proj.bias']
    x = layer_norm(x + sa, sd[prefix + 'norm1.weight'], sd[prefix + 'norm1.bias'])
    h = torch.relu(x @ sd[prefix + 'linear1.weight'].T + sd[prefix + 'linear1.bias'])
    ff = h @ sd[prefix + 'linear2.weight'].T + sd[prefix + 'linear2.bias']
    return layer_norm(x + ff, sd[prefix + 'norm2.weight'], sd[prefix + 'norm2.bias'])


def transformer_body(sd, x, lengths, heads=ATTENTION_HEADS, is_causal=False):
    """ppgs/model/transformer.py:65-81 for one (un-chunked) padded batch.
    x (B,C,T) with T == max(lengths); returns logits (B,40,T)."""
    layers = 1 + max(
        int(key.split('.')[2]) for key in sd if key.startswith('model.layers.'))
    mask = mask_from_lengths(lengths)                        # (B,T)
    if mask.shape[1] != x.shape[-1]:
        raise ValueError('max(lengths) must equal the padded length')
    h = conv1d_same(x, sd['input_layer.weight'], sd['input_layer.bias'])
    h = h * mask[:, None, :].to(h.dtype)
    T = x.shape[-1]
    if T > sd['position.encoding'].shape[0]:
        raise ValueError('size is too large')
    h = h.permute(0, 2, 1) + sd['position.encoding'][:T, 0][None]
    for layer in range(layers):
        h = encoder_layer(h, sd, f'model.layers.{layer}.', heads, mask, is_causal)
    out = conv1d_same(h.permute(0, 2, 1), sd['output_layer.weight'], sd['output_layer.bias'])
    return out * mask[:, None, :].to(out.dtype)


def chunk_plan(T, lengths):
    """The length bookkeeping of ppgs/model/transformer.py:49-64.

    Returns a list of (start, stop, chunk_lengths) over the left-padded axis
    (`padded = replicate-pad(x, (50, 0))`), one entry per block.
    """
    stride = CHUNK_LENGTH - 2 * CHUNK_OVERLAP
    lengths = lengths.clone()
    plan = []
    for i in range(math.ceil(T / stride)):
        start = i * stride
        stop = min((i + 1) * stride + 2 * CHUNK_OVERLAP, T + CHUNK_OVERLAP)
        chunk_lengths = (lengths + CHUNK_OVERLAP).clamp(0, CHUNK_LENGTH)
        chunk_lengths[chunk_lengths == CHUNK_OVERLAP] = 0
        lengths = (lengths - stride).clamp(min=0)
        plan.append((start, stop, chunk_lengths))
    return plan


def transformer_forward(sd, x, lengths, heads=ATTENTION_HEADS, is_causal=False,
                        legacy_mode=False):
    """ppgs/model/transformer.py:45-81 including the 500/400/50 chunking."""
    x = x.to(sd['input_layer.weight'].dtype)
    T = x.shape[-1]
    if legacy_mode:
        assert T < MAX_LEN
    elif T > CHUNK_LENGTH:
        padded = torch.nn.functional.pad(x, (CHUNK_OVERLAP, 0), mode='replicate')
        outs = []
        for start, stop, chunk_lengths in chunk_plan(T, lengths):
            split = padded[..., start:stop]
            if int(chunk_lengths.max()) != split.shape[-1]:
                # reference would fail at `* mask` (SURVEY.md §3.2 i)
                raise ValueError('max(lengths) must equal the padded length')
            out = transformer_body(sd, split, chunk_lengths, heads, is_causal)
            outs.append(out[..., CHUNK_OVERLAP:CHUNK_LENGTH - CHUNK_OVERLAP])
        return torch.cat(outs, dim=-1)
    return transformer_body(sd, x, lengths, heads, is_causal)


def cast_state_dict(sd, dtype):
    return {k: v.to(dtype) for k, v in sd.items()}


def from_features(sd, features, lengths, softmax=True, is_causal=False,
                  legacy_mode=False, dtype=torch.float32,
                  heads=ATTENTION_HEADS):
    """ppgs/core.py:72-128,551-596 with autocast disabled (oracle mode O3)."""
    sd = cast_state_dict(sd, dtype)
    logits = transformer_forward(
        sd, features.to(dtype), lengths, heads, is_causal, legacy_mode)
    if softmax:
        return torch.softmax(logits, dim=1)
    return logits


def from_audio(sd, audio, softmax=True, is_causal=False, legacy_mode=False,
               dtype=torch.float32, lengths=None):
    """Batched ppgs.from_audio (ppgs/core.py:22-69) == mel.from_audios +
    from_features (ppgs/core.py:333-352), O3 numerics.  audio (B,1,samples)."""
    features = mel_from_audios(audio)
    if lengths is None:
        lengths = torch.full((audio.shape[0],), features.shape[-1], dtype=torch.long)
    else:
        lengths = lengths // HOPSIZE
    return from_features(sd, features, lengths, softmax, is_causal, legacy_mode, dtype)


###############################################################################
# Batch scheduler / file path restatements
###############################################################################


def sampler_batches(frame_lengths, max_frames, seed=1234, epoch=0):
    """ppgs/data/sampler.py:46-82 + ppgs/data/dataset.py:107-127 with BUCKETS=1
    (ppgs/config/defaults.py:170) and RANDOM_SEED=1234 (:202).

    One bucket = all indices in argsort(length) order; shuffled with
    randperm(seed+epoch); greedily packed while (n+1)*max_len <= max_frames;
    the list of batches is shuffled again with the same generator.  Files longer
    than max_frames are dropped by Metadata (ppgs/data/dataset.py:189-198)
    before this is called.
    """
    g = torch.Generator()
    g.manual_seed(seed + epoch)
    lengths = np.asarray(frame_lengths)
    indices = np.argsort(lengths)
    bucket = np.stack((indices, np.sort(lengths))).T
    bucket = bucket[torch.randperm(len(bucket), generator=g).tolist()]
    batches, batch, max_length = [], [], 0
    for index, length in bucket:
        max_length = max(max_length, length)
        if batch and (len(batch) + 1) * max_length > max_frames:
            batches.append(batch)
            max_length = length
            batch = [int(index)]
        else:
            batch.append(int(index))
    if batch:
        batches.append(batch)
    return [batches[i] for i in torch.randperm(len(batches), generator=g).tolist()]


def collate(audios):
    """ppgs/data/collate.py:19-28: zero-pad (1,samples_i) audios to
    (B,1,max_samples) fp32 + int64 lengths."""
    lengths = torch.tensor([a.shape[-1] for a in audios], dtype=torch.long)
    padded = torch.zeros(len(audios), 1, int(lengths.max()), dtype=torch.float32)
    for i, a in enumerate(audios):
        padded[i, :, :a.shape[-1]] = a
    return padded, lengths


def save_masked_crop(tensor, length):
    """ppgs/preprocess/core.py:219-221."""
    return tensor[..., :length].clone()


###############################################################################
# As-shipped module stack (CPU baseline timing only)
###############################################################################


class AsShipped:
    """The reference's CPU path as shipped (oracle mode O1): stock torch layers
    (`Conv1d`, `nn.TransformerEncoder`, `torch.stft`) under
    `torch.autocast('cpu')` + inference_mode, i.e. bf16 GEMMs
    (ppgs/core.py:586, ppgs/preprocess/core.py:207).  Used by
    `bench.py --impl reference` and `cpu_baseline`; the reference tree itself
    cannot travel to the GPU box.
    """

    def __init__(self, sd, is_causal=False):
        import warnings
        H, C, k = sd['input_layer.weight'].shape
        O = sd['output_layer.weight'].shape[0]
        layers = 1 + max(int(key.split('.')[2]) for key in sd
                         if key.startswith('model.layers.'))
        self.input_layer = torch.nn.Conv1d(C, H, k, padding='same')
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            self.model = torch.nn.TransformerEncoder(
                torch.nn.TransformerEncoderLayer(H, ATTENTION_HEADS), layers)
        self.output_layer = torch.nn.Conv1d(H, O, k, padding='same')
        self.encoding = sd['position.encoding']
        self.is_causal = is_causal
        self.input_layer.load_state_dict(
            {'weight': sd['input_layer.weight'], 'bias': sd['input_layer.bias']})
        self.output_layer.load_state_dict(
            {'weight': sd['output_layer.weight'], 'bias': sd['output_layer.bias']})
        self.model.load_state_dict(
            {k[len('model.'):]: v for k, v in sd.items() if k.startswith('model.')})
        for m in (self.input_layer, self.model, self.output_layer):
            m.eval()
        self.basis = torch.from_numpy(mel_basis())

    def body(self, x, lengths):
        causal_mask = None
        if self.is_causal:
            causal_mask = torch.nn.Transformer.generate_square_subsequent_mask(
                int(lengths.max()))
        mask = mask_from_lengths(lengths).unsqueeze(1)
        x = self.input_layer(x) * mask
        x = x.permute(2, 0, 1)
        x = x + self.encoding[:x.size(0)]
        x = self.model(x, mask=causal_mask,
                       src_key_padding_mask=~mask.squeeze(1)).permute(1, 2, 0)
        return self.output_layer(x) * mask

    def forward(self, x, lengths):
        T = x.shape[-1]
        if T > CHUNK_LENGTH:
            padded = torch.nn.functional.pad(x, (CHUNK_OVERLAP, 0), mode='replicate')
            outs = []
            for start, stop, chunk_lengths in chunk_plan(T, lengths):
                out = self.body(padded[..., start:stop], chunk_lengths)
                outs.append(out[..., CHUNK_OVERLAP:CHUNK_LENGTH - CHUNK_OVERLAP])
            return torch.cat(outs, dim=-1)
        return self.body(x, lengths)

    def from_audios(self, audio, autocast=True):
        """audio (B,1,samples) fp32 -> posteriors (B,40,frames)."""
        with torch.inference_mode():
            features = mel_from_audios(audio)   # dataloader path: no autocast
            lengths = torch.full(
                (audio.shape[0],), features.shape[-1], dtype=torch.long)
            with torch.autocast('cpu', enabled=autocast):
                logits = self.forward(features, lengths)
                return torch.softmax(logits, dim=1)
