/*
 * ppgs_b200 — C ABI of the B200-native PPG inference engine.
 *
 * Drop-in boundary for the forward path of interactiveaudiolab/ppgs
 * (`ppgs.from_audio`: mel front-end -> Transformer encoder -> 40-way softmax).
 * The reference has no native code and no FFI; each entry point
 * below names the reference Python function whose arithmetic it replaces, and
 * `INTEGRATION.md` shows the ctypes binding a maintainer of the reference would
 * add at that call site.
 *
 * Conventions
 *  - plain C types only; all `*_dev` pointers are CUDA device pointers on the
 *    engine's device, all `*_host` pointers are host memory;
 *  - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default
 *    stream); device entry points are asynchronous on that stream;
 *  - every function returns 0 on success or a negative `PPGS_E_*` code and never
 *    aborts; `ppgs_last_error()` returns a thread-local message for the last
 *    failure;
 *  - one engine per (device, checkpoint); calls on one engine must be serialised
 *    by the caller (the reference's function-attribute caches are not
 *    thread-safe either: ppgs/core.py:565);
 *  - the engine owns the packed weights and a grow-only device workspace; the
 *    caller owns every input / output buffer.
 */
#ifndef PPGS_B200_H_
#define PPGS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPGS_ABI_VERSION 1

/* error codes (mirrors of the reference's exception types are noted) */
#define PPGS_OK 0
#define PPGS_E_INVALID -1    /* ValueError: bad argument / shape / unknown name */
#define PPGS_E_CUDA -2       /* RuntimeError: a CUDA call failed               */
#define PPGS_E_STATE -3      /* RuntimeError: engine not finalised / missing weight */
#define PPGS_E_TOO_LARGE -4  /* ValueError('size is too large'): transformer.py:103-104 */
#define PPGS_E_UNSUPPORTED -5

/* arithmetic of the dense contractions (GEMMs + attention) */
#define PPGS_PRECISION_FP32 0       /* CUDA-core FFMA; validation mode               */
#define PPGS_PRECISION_F16X2 1      /* tcgen05, split-fp16 operands, 3 MMA passes, fp32
                                       accumulate in TMEM: the <=1e-4 parity mode    */
#define PPGS_PRECISION_F16 2        /* tcgen05, single-pass fp16 operands (the
                                       reference's own CUDA autocast numerics class) */

/* Model hyper-parameters: ppgs/config/defaults.py:127-161, ppgs/config/w2v2fb.py:7-10,
 * config/causal_transformer.py:18, torch.nn.TransformerEncoderLayer defaults. */
typedef struct ppgs_model_config {
    int32_t input_channels;   /* 80 (mel) / 768 (w2v2fb)  */
    int32_t hidden_channels;  /* 256 / 512                */
    int32_t num_layers;       /* 5                        */
    int32_t num_heads;        /* 2                        */
    int32_t ffn_channels;     /* 2048                     */
    int32_t output_channels;  /* 40 = len(ppgs.PHONEMES)  */
    int32_t kernel_size;      /* 5                        */
    int32_t is_causal;        /* IS_CAUSAL                */
    int32_t chunk_length;     /* 500                      */
    int32_t chunk_overlap;    /* 50                       */
    int32_t max_len;          /* 5000 (positional table)  */
    float layer_norm_eps;     /* 1e-5                     */
} ppgs_model_config;

typedef struct ppgs_engine ppgs_engine;

int ppgs_abi_version(void);
const char* ppgs_last_error(void);

/* Fills *cfg with the reference defaults for 'mel' (ppgs/config/defaults.py). */
void ppgs_default_config(ppgs_model_config* cfg);

/* ppgs.load.model (ppgs/load.py:33-81) + ppgs.Model (ppgs/model/core.py:9-25):
 * build an engine for `cfg` on CUDA device `device`. */
int ppgs_engine_create(const ppgs_model_config* cfg, int device, ppgs_engine** out);
void ppgs_engine_destroy(ppgs_engine* engine);

/* `model.load_state_dict` (ppgs/load.py:76-79): upload one fp32 tensor under its
 * reference state-dict key, e.g. "input_layer.weight" (H,C,5),
 * "model.layers.0.self_attn.in_proj_weight" (3H,H), "position.encoding" (L,1,H).
 * Shapes are validated against `cfg`; unknown keys are an error (strict load). */
int ppgs_engine_set_weight(ppgs_engine* engine, const char* name,
                           const float* data_host, const int64_t* shape, int ndim);

/* Strict-load check (every key present) + packing into kernel layouts (k-major
 * conv weights, split-fp16 planes, power-of-two scales). */
int ppgs_engine_finalize(ppgs_engine* engine);

/* Default (chosen at finalize): PPGS_PRECISION_F16X2 when the model shape has tensor-core
 * kernels (hidden 256, head_dim 128 — the mel model), else PPGS_PRECISION_FP32. */
int ppgs_engine_set_precision(ppgs_engine* engine, int precision);
int ppgs_engine_get_precision(const ppgs_engine* engine);

/* Single NCCL-free hook for multi-GPU loading: size of / pointer to the packed
 * weight blob on the device, so the host can `torch.distributed.broadcast` it
 * from rank 0 before `ppgs_engine_adopt_blob`. */
size_t ppgs_engine_blob_bytes(const ppgs_engine* engine);
void* ppgs_engine_blob_dev(ppgs_engine* engine);
/* After the blob bytes were overwritten in place (e.g. by a broadcast), mark the
 * engine finalised without re-uploading from host tensors. */
int ppgs_engine_adopt_blob(ppgs_engine* engine);

/* ppgs.preprocess.mel.from_audios (ppgs/preprocess/mel.py:14-19) =
 * spectrogram.from_audios (ppgs/preprocess/spectrogram.py:14-50) + linear_to_mel
 * (mel.py:56-76), autocast off.
 *   audio_dev : (batch, samples) fp32, row stride `audio_stride` elements
 *   mel_dev   : (batch, 80, samples/160) fp16, contiguous
 * Requires samples >= 433 (reflect padding of 432, as torch does). */
int ppgs_mel_forward(ppgs_engine* engine, const float* audio_dev, int batch,
                     int64_t samples, int64_t audio_stride, void* mel_dev,
                     void* stream);

/* ppgs.preprocess.w2v2fb.from_audios (ppgs/preprocess/w2v2fb/core.py:32-75): the
 * Hugging Face Wav2Vec2Model('facebook/wav2vec2-base') forward on the audio padded by 40
 * zeros each side, with the padding mask derived from `lengths_host` (samples; NULL = all
 * full length), `last_hidden_state` nearest-upsampled to samples/160 frames, fp16.
 * The wav2vec2 weights are uploaded with ppgs_engine_set_weight under their Hugging Face
 * state-dict keys prefixed by "w2v2." (e.g. "w2v2.encoder.layers.0.attention.q_proj.weight"),
 * then packed by ppgs_w2v2_finalize (strict: every key, exact shapes; folds the weight
 * norm of the positional convolution).
 *   audio_dev    : (batch, samples) fp32, row stride `audio_stride`
 *   features_dev : (batch, 768, samples/160) fp16, contiguous
 * Encoder projections / FFN run as split-fp16 tcgen05 GEMMs, everything else of the
 * front-end in fp32 on the CUDA cores (PPGS_B200_W2V2_TC=0: all fp32). */
int ppgs_w2v2_finalize(ppgs_engine* engine);
int ppgs_w2v2fb_forward(ppgs_engine* engine, const float* audio_dev, int batch, int64_t samples,
                        int64_t audio_stride, const int64_t* lengths_host, void* features_dev,
                        void* stream);

/* ppgs.from_features -> ppgs.infer -> Transformer.forward (ppgs/core.py:72-128,
 * 551-596; ppgs/model/transformer.py:45-81), including the 500/400/50 chunking
 * unless `legacy_mode`, the key-padding (and causal) masks, and softmax(dim=1)
 * when `softmax` != 0.
 *   features_dev : (batch, input_channels, frames) fp16, contiguous
 *   lengths_host : (batch,) int64 frame lengths; max must equal `frames`
 *   out_dev      : (batch, output_channels, frames) fp32, contiguous */
int ppgs_transformer_forward(ppgs_engine* engine, const void* features_dev,
                             int batch, int frames, const int64_t* lengths_host,
                             int softmax, int legacy_mode, float* out_dev,
                             void* stream);

/* ppgs.from_audio (ppgs/core.py:22-69) for a batch, device buffers:
 * mel front-end + transformer in one call; the fp16 features stay in the
 * engine workspace.  lengths_host are SAMPLE lengths (NULL = all `samples`),
 * converted with `// 160` like ppgs/core.py:326. */
int ppgs_from_audio(ppgs_engine* engine, const float* audio_dev, int batch,
                    int64_t samples, int64_t audio_stride,
                    const int64_t* lengths_host, int softmax, int legacy_mode,
                    float* out_dev, void* stream);

/* Same, HOST buffers (the reference-facing call timed as `e2e` in bench.py):
 * H2D of the audio, compute, D2H of the posteriors, stream-synchronised on
 * return.  Pinned host memory is used as-is; pageable memory works but is slower. */
int ppgs_from_audio_host(ppgs_engine* engine, const float* audio_host, int batch,
                         int64_t samples, const int64_t* lengths_host,
                         int softmax, int legacy_mode, float* out_host,
                         void* stream);

/* Pipelined form of `ppgs_from_audio_host` for serving / file loops (the batching
 * loop of ppgs/core.py:325-365): enqueue the request and return.  Two requests can be
 * in flight per engine — the H2D copy of request i+1 and the D2H copy of request
 * i-1 run on the engine's copy streams while the kernels of request i run on
 * `stream`.  A third submit blocks until the oldest request has landed.  `out_host`
 * is valid after `ppgs_engine_wait`, which blocks until every submitted request
 * has completed.  Host buffers must stay alive (and should be pinned) until then. */
int ppgs_from_audio_host_submit(ppgs_engine* engine, const float* audio_host, int batch,
                                int64_t samples, const int64_t* lengths_host, int softmax,
                                int legacy_mode, float* out_host, void* stream);
int ppgs_engine_wait(ppgs_engine* engine);

/* Number of kernels this library launched since the engine was created
 * (bench.py's `gpu_launches`). */
int64_t ppgs_engine_launch_count(const ppgs_engine* engine);

/* Steady-state loops of ppgs_from_audio / ppgs_from_audio_host* / ppgs_files_to_files (same
 * shapes, lengths and buffers as an earlier call, engine workspace untouched in between) are
 * replayed as one CUDA graph instead of 29 launches.  Kernels inside a replay still count in
 * ppgs_engine_launch_count.  On by default (tensor-core precisions, profiling off);
 * PPGS_B200_GRAPHS=0 or ppgs_engine_set_graphs(engine, 0) keeps the launch path. */
int64_t ppgs_engine_graph_replays(const ppgs_engine* engine);
int ppgs_engine_set_graphs(ppgs_engine* engine, int enabled);

/* Per-kernel device timing for bench.py's roofline: when enabled, every launch is
 * bracketed by CUDA events on its stream and accumulated under the kernel's
 * name.  `ppgs_engine_kernel_stat(e, i, ...)` enumerates the names (returns
 * PPGS_E_INVALID past the end; synchronises the pending events).  Enabling or
 * disabling clears the statistics. */
int ppgs_engine_set_profiling(ppgs_engine* engine, int enabled);
int ppgs_engine_kernel_stat(ppgs_engine* engine, int index, char* name, size_t name_bytes,
                            double* total_ms, int64_t* launches);

/* Device bytes currently held by the engine's grow-only workspace. */
size_t ppgs_engine_workspace_bytes(const ppgs_engine* engine);

/* ---- audio ingest / posteriorgram egress of the file API -------------------------------- */

/* torchaudio.info as used by ppgs/data/dataset.py:187 (list-of-files metadata): header
 * probe of a RIFF/WAVE file (PCM 8/16/24/32 bit, IEEE float 32/64, WAVE_FORMAT_EXTENSIBLE).
 * Any out pointer may be NULL.  PPGS_E_UNSUPPORTED: not a WAVE file / compressed codec. */
int ppgs_wav_info(const char* path, int64_t* frames, int* sample_rate, int* channels,
                  int* bits, int* is_float);

/* The same probe over a file list on `threads` host threads (the Metadata loop of
 * ppgs/data/dataset.py:180-198 for a corpus); status[i] is the per-file PPGS_* code. */
int ppgs_wav_info_many(const char* const* paths, int64_t count, int threads, int64_t* frames,
                       int32_t* sample_rate, int32_t* channels, int32_t* bits, int32_t* is_float,
                       int32_t* status);

/* torchaudio.load as used by ppgs/load.py:17-30, for WAVE files: channel 0 (what
 * ppgs/data/collate.py:27 keeps) as fp32 normalised to [-1, 1) (int16 / 32768 ...). */
int ppgs_wav_read_f32(const char* path, float* dst_host, int64_t capacity, int64_t* frames,
                      int* sample_rate);

/* torchaudio.info / torchaudio.load (ppgs/data/dataset.py:187, ppgs/load.py:17-30) for FLAC
 * files: STREAMINFO probe, and a verifying decode (frame CRC-8 / CRC-16, STREAMINFO MD5) of ALL
 * channels to fp32 = sample / 2^(bits-1), channel-major with row stride `capacity` frames.
 * PPGS_E_UNSUPPORTED: not a FLAC stream.  Any out pointer may be NULL. */
int ppgs_flac_info(const char* path, int64_t* frames, int* sample_rate, int* channels, int* bits);
int ppgs_flac_read_f32(const char* path, float* dst_host, int64_t capacity, int64_t* frames,
                       int* sample_rate, int* channels);

/* The /32768 normalisation of torchaudio.load on the device: `count` int16 samples ->
 * fp32 (both buffers 16-byte aligned), so that file batches cross PCIe as 2-byte PCM. */
int ppgs_pcm16_to_f32(ppgs_engine* engine, const void* pcm_dev, int64_t count,
                      float* out_dev, void* stream);

/* ppgs.resample (ppgs/core.py:599-608) = torchaudio.transforms.Resample(orig, target)
 * with its defaults (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99) on the
 * device.  Output length per row = ppgs_resample_length = ceil(target * samples / orig).
 *   audio_dev : (batch, samples) fp32, row stride `audio_stride`
 *   out_dev   : (batch, out_len) fp32, row stride `out_stride` */
int64_t ppgs_resample_length(int64_t samples, int orig_rate, int target_rate);
int ppgs_resample(ppgs_engine* engine, const float* audio_dev, int batch, int64_t samples,
                  int64_t audio_stride, int orig_rate, int target_rate, float* out_dev,
                  int64_t out_stride, void* stream);
/* The filter table of that resampler (host; for inspection / tests): `taps` receives
 * [ntaps][phases] fp32 when not NULL; phases = target/gcd, ntaps = 2*width + orig/gcd. */
int ppgs_resample_taps(int orig_rate, int target_rate, float* taps, int64_t capacity,
                       int* ntaps, int* phases, int* width);

/* torch.save(tensor) of ppgs.preprocess.save_masked (ppgs/preprocess/core.py:219-221)
 * without Python: writes a torch.load-compatible archive holding the contiguous fp32
 * tensor (rows, cols) read from `data` with row stride `row_stride` (the crop of a padded
 * batch row is a strided read). */
int ppgs_pt_write_f32(const char* path, const float* data_host, int64_t rows, int64_t cols,
                      int64_t row_stride);
/* Same for an fp16 tensor (torch.HalfStorage): the feature files of `python -m ppgs.preprocess`
 * (ppgs/preprocess/core.py:105-190 saves the fp16 representations with save_masked). */
int ppgs_pt_write_f16(const char* path, const void* data_host, int64_t rows, int64_t cols,
                      int64_t row_stride);

/* torch.load of a feature cache (ppgs/data/dataset.py:98-101: `torch.load(cache /
 * f'{stem}-{feature}.pt')`) without Python: probe / read a torch.save archive that holds ONE
 * whole contiguous fp16 or fp32 CPU tensor of 1..3 dimensions (what ppgs.preprocess and this
 * library write).  ppgs_pt_info: dims3 = the sizes right-aligned in 3 slots (leading slots 1).
 * ppgs_pt_read: copies the tensor viewed as (rows, cols) = (dims3[0]*dims3[1], dims3[2]) into
 * `dst_host` with row stride `dst_row_stride` elements (a row of a padded batch).  Archives the
 * native reader does not understand (legacy format, zip64, other dtypes, views) return
 * PPGS_E_UNSUPPORTED: callers fall back to torch.load. */
int ppgs_pt_info(const char* path, int* ndim, int64_t* dims3, int* elem_bytes);
int ppgs_pt_read(const char* path, void* dst_host, int64_t rows, int64_t cols, int elem_bytes,
                 int64_t dst_row_stride);

/* The batching loop of ppgs.from_files_to_files / from_dataloader (ppgs/core.py:207-391)
 * for the mel representation as ONE call: `reader_threads` threads decode 16-bit PCM
 * 16 kHz WAVE files (and 16 kHz FLAC files of <= 16 bits, by extension `.flac`, verified
 * like ppgs_flac_read_f32) straight into pinned zero-padded int16 batches, the calling thread
 * runs H2D -> pcm16_to_f32 -> mel + Transformer + softmax -> D2H on two device slots
 * (copies on private streams, kernels on `stream`), `writer_threads` threads crop every
 * row to samples/160 frames and write `<output>.pt`.  The threads live for the duration
 * of the call only.
 *   batch_sizes  : files per batch, n_batches entries (batch composition is the
 *                  caller's: ppgs/data/sampler.py:46-82 semantics live in Python)
 *   audio_files / output_files / file_samples : flat, in batch order; file_samples are
 *                  the per-file sample counts from ppgs_wav_info / ppgs_flac_info
 *   frames_done  : out, posteriorgram frames written
 * PPGS_E_UNSUPPORTED when a file is not 16-bit PCM at 16 kHz (callers fall back to the
 * per-batch API with ppgs_wav_read_f32 + ppgs_resample). */
int ppgs_files_to_files(ppgs_engine* engine, int n_batches, const int32_t* batch_sizes,
                        const char* const* audio_files, const char* const* output_files,
                        const int64_t* file_samples, int reader_threads, int writer_threads,
                        int legacy_mode, void* stream, int64_t* frames_done);

/* ---- stateful streaming decoder for causal models (config/causal_transformer.py:18) ------
 * The reference has no streaming state (ppgs/model/transformer.py:65-71 rebuilds the causal
 * mask per call); this is the incremental form of its un-chunked causal forward
 * (`legacy_mode=True`, IS_CAUSAL=True) over `streams` independent utterances.
 * State per stream: the feature rows, and per layer the K / V rows of every frame pushed so
 * far (the attention cache, read in place by the tcgen05 attention kernel).  A session holds
 * ppgs_stream_capacity() = 510 frames; PPGS_E_TOO_LARGE beyond (ValueError('size is too
 * large'), transformer.py:103-104) — restart with overlap like the reference's chunking.
 *
 * ppgs_stream_push appends `frames` feature frames (features_dev: (streams, input_channels,
 * frames) fp16, contiguous) and writes the posteriorgram frames that became final — frames
 * [emitted, length - 4): 2 + 2 frames of look-ahead of the two k=5 convolutions — to out_dev
 * ((streams, output_channels, out_capacity) fp32; the first *frames_out frames of every row
 * are valid).  `final` != 0 also emits the last 4 frames (zero padding beyond the end, as the
 * reference pads) and closes the session until ppgs_stream_reset.  Every emitted frame equals
 * the reference's value for the whole utterance. */
typedef struct ppgs_stream ppgs_stream;
int ppgs_stream_capacity(void);
int ppgs_stream_create(ppgs_engine* engine, int streams, ppgs_stream** out);
void ppgs_stream_destroy(ppgs_stream* stream_state);
int ppgs_stream_reset(ppgs_stream* stream_state, void* stream);
int ppgs_stream_length(const ppgs_stream* stream_state);
int ppgs_stream_emitted(const ppgs_stream* stream_state);
int ppgs_stream_push(ppgs_stream* stream_state, const void* features_dev, int frames, int final,
                     int softmax, float* out_dev, int out_capacity, int* frames_out,
                     void* stream);
/* Independent streams (serving): stream b brings frames[b] <= max_frames frames (the first
 * frames[b] columns of row b of features_dev, (streams, input_channels, max_frames) fp16), may
 * be finalised on its own (final[b], NULL = none) and emits frames_out[b] frames into row b of
 * out_dev.  ppgs_stream_reset_streams clears the flagged streams for their next utterance
 * while the others keep their state; ppgs_stream_state reports per-stream lengths / emitted
 * frames (either pointer may be NULL).  ppgs_stream_push is the lockstep special case. */
int ppgs_stream_push_ragged(ppgs_stream* stream_state, const void* features_dev, int max_frames,
                            const int32_t* frames, const int32_t* final, int softmax,
                            float* out_dev, int out_capacity, int32_t* frames_out, void* stream);
int ppgs_stream_reset_streams(ppgs_stream* stream_state, const int32_t* flags, void* stream);
int ppgs_stream_state(const ppgs_stream* stream_state, int32_t* lengths, int32_t* emitted);

/* ---- posteriorgram post-processing on the device (the step after the hot path) ----------- */

/* ppgs.distance (ppgs/core.py:399-469): x, y = (phonemes, frames) fp32 with row strides;
 * similarity_dev = (phonemes, phonemes) row-major similarity matrix (the reference's
 * assets/balanced_similarity.pt) or NULL for normalize=False; reduction 0 'none' (out:
 * frames floats), 1 'mean', 2 'sum' (out: one float). */
int ppgs_ppg_distance(ppgs_engine* engine, const float* x_dev, const float* y_dev, int phonemes,
                      int64_t frames, int64_t x_stride, int64_t y_stride,
                      const float* similarity_dev, float exponent, int reduction, float* out_dev,
                      void* stream);
/* ppgs.interpolate (ppgs/core.py:475-496): (1 - w) x + w y over (rows, frames) contiguous;
 * interp_dev = per-frame weights or NULL for the scalar. */
int ppgs_ppg_interpolate(ppgs_engine* engine, const float* x_dev, const float* y_dev,
                         const float* interp_dev, float interp_scalar, int64_t rows,
                         int64_t frames, float* out_dev, void* stream);
/* ppgs.edit.grid.sample (ppgs/edit/grid.py:13-50): ppg (phonemes, frames) contiguous, grid of
 * float frame indices -> out (phonemes, grid_len). */
int ppgs_ppg_grid_sample(ppgs_engine* engine, const float* ppg_dev, int phonemes, int64_t frames,
                         const float* grid_dev, int64_t grid_len, float* out_dev, void* stream);
/* ppgs.sparsify (ppgs/core.py:504-543): ppg (batch, phonemes, frames) contiguous; method 0
 * 'constant', 1 'percentile' (threshold = q), 2 'topk' (threshold = k); renormalised with
 * softmax(log(ppg + 1e-8)). */
int ppgs_ppg_sparsify(ppgs_engine* engine, const float* ppg_dev, int batch, int phonemes,
                      int64_t frames, int method, float threshold, float* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PPGS_B200_H_ */
