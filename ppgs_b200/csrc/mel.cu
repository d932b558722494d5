// K1: fused framed STFT + Slaney mel + log + fp16 kernel (HBM-bound stage).
//
// Replaces ppgs.preprocess.mel.from_audios (ppgs/preprocess/mel.py:14-19):
//   spectrogram.from_audios  ppgs/preprocess/spectrogram.py:14-50
//   linear_to_mel            ppgs/preprocess/mel.py:56-76
// which in the reference is reflection_pad1d + as_strided framing + cuFFT R2C +
// five elementwise passes over a complex (B,513,T) tensor + a batched GEMM +
// clamp/log/half.  Here one persistent CTA stages the audio of a 32-frame tile
// in shared memory with the bulk async-copy engine (cp.async.bulk -> UBLKCP,
// mbarrier completion), each warp runs 1024-point real FFTs out of that tile and
// only the (80, frames) fp16 mel leaves the SM: 640 B in + 160 B out per frame.
//
// Compiled with --fmad=false: the arithmetic is written with explicit fmaf so
// that the CPU emulation harness (tests/csrc/mel_emul.cpp) replays it bit for bit.
#include <math.h>

#include "common.cuh"
#include "mel_math.cuh"

namespace ppgs {

constexpr int kTileFrames = 32;
constexpr int kMelWarps = 8;
constexpr int kMelThreads = kMelWarps * 32;
constexpr int kTileSamples = (kTileFrames - 1) * kHop + kNfft;   // 5984
constexpr int kRowPitch = 88;      // halves per frame row of the transposed output tile

struct MelSmem {
    alignas(16) float audio[kTileSamples];
    alignas(16) float window[kNfft];
    alignas(16) cf tw_pass1[7 * 8];     // [r - 1][k]      = tw512[8 k r]
    alignas(16) cf tw_pass2[7 * 64];    // [r - 1][k]      = tw512[k r]
    alignas(16) cf tw1024[kBins + 3];
    alignas(16) cf zb[kMelWarps][kZPad];
    alignas(16) float fb_w[kFbMaxRounds * kFbPiece * 32];
    int32_t fb_base[kFbMaxRounds * 32];
    int32_t band_slot[kMels], band_pieces[kMels];
    // (B, 80, T) output: rows padded to 34 halves (68 B) so that the per-frame column store
    // out[m][f] of lanes m = 0..31 hits 32 different banks (17 m mod 32) instead of two.
    // Operand-row output: out_t[f][m], rows of 88 halves (176 B, 16-byte segments for the stores).
    union {
        alignas(16) __half out[kMels][kTileFrames + 2];
        alignas(16) __half out_t[kTileFrames][kRowPitch];
    };
    alignas(8) unsigned long long mbar;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes,
                                              unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// One frame by one warp.  `a` points at the frame's first (padded-signal) sample
// in the shared audio tile.
template <bool ROWS>
__device__ __forceinline__ void frame_to_mel(const float* a, MelSmem& s, cf* zb, int lane,
                                             int frame_in_tile, int fb_rounds) {
    cf v0[8], v1[8];
    // ---- pass 0: window + pack (z[n] = x[2n] + i x[2n+1]) + radix-8, Ns = 1
    {
        const float2* a2 = reinterpret_cast<const float2*>(a);
        const float2* w2 = reinterpret_cast<const float2*>(s.window);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float2 x = a2[lane + 64 * r], w = w2[lane + 64 * r];
            v0[r] = {x.x * w.x, x.y * w.y};
            x = a2[lane + 32 + 64 * r];
            w = w2[lane + 32 + 64 * r];
            v1[r] = {x.x * w.x, x.y * w.y};
        }
        dft8(v0);
        dft8(v1);
        __syncwarp();   // previous frame's readers of zb are done
        const int b0 = stockham_store_base<0>(lane), b1 = stockham_store_base<0>(lane + 32);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            zb[zpad(b0 + r)] = v0[r];
            zb[zpad(b1 + r)] = v1[r];
        }
    }
    // ---- pass 1: Ns = 8
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        v0[r] = zb[zpad(lane + 64 * r)];
        v1[r] = zb[zpad(lane + 32 + 64 * r)];
    }
    stockham_twiddle_table<1>(v0, lane, s.tw_pass1);
    stockham_twiddle_table<1>(v1, lane + 32, s.tw_pass1);
    dft8(v0);
    dft8(v1);
    __syncwarp();
    {
        const int b0 = stockham_store_base<1>(lane), b1 = stockham_store_base<1>(lane + 32);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            zb[zpad(b0 + 8 * r)] = v0[r];
            zb[zpad(b1 + 8 * r)] = v1[r];
        }
    }
    // ---- pass 2: Ns = 64
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        v0[r] = zb[zpad(lane + 64 * r)];
        v1[r] = zb[zpad(lane + 32 + 64 * r)];
    }
    stockham_twiddle_table<2>(v0, lane, s.tw_pass2);
    stockham_twiddle_table<2>(v1, lane + 32, s.tw_pass2);
    dft8(v0);
    dft8(v1);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        zb[zpad(lane + 64 * r)] = v0[r];
        zb[zpad(lane + 32 + 64 * r)] = v1[r];
    }
    __syncwarp();
    // ---- unpack to 513 one-sided bins, magnitude, fp16 round: bins k and 512 - k share one
    // pair of FFT outputs (lane 0 also takes the three self-paired bins 0, 256 and 512)
    float mlo[8], mhi[8], mid = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = lane + 32 * i;
        float plo, phi;
        if (k == 0) {
            plo = bin_power(zb, 0, s.tw1024);
            phi = bin_power(zb, kHalf, s.tw1024);
            mid = __half2float(__float2half_rn(sqrtf(bin_power(zb, kHalf / 2, s.tw1024))));
        } else {
            bin_power_pair(zb, k, s.tw1024, plo, phi);
        }
        mlo[i] = __half2float(__float2half_rn(sqrtf(plo)));
        mhi[i] = __half2float(__float2half_rn(sqrtf(phi)));
    }
    __syncwarp();
    float* spec = reinterpret_cast<float*>(zb);   // reuse the FFT buffer: [513] floats
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = lane + 32 * i;
        spec[k] = mlo[i];
        spec[kHalf - k] = mhi[i];
    }
    if (lane == 0) spec[kHalf / 2] = mid;
    __syncwarp();
    // ---- triangular filterbank: one piece of kFbPiece bins per lane and round, then the
    // bands add their pieces' partial sums (mel_math.cuh)
    float* partial = spec + kFbPartialOffset;
    for (int r = 0; r < fb_rounds; ++r)
        partial[r * 32 + lane] =
            filterbank_piece(s.fb_w + r * kFbPiece * 32, s.fb_base + r * 32, lane, spec);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int m = lane + 32 * i;
        if (m < kMels) {
            const float acc = filterbank_band(partial, s.band_slot[m], s.band_pieces[m]);
            const __half v = __float2half_rn(logf(fmaxf(acc, 1e-5f)));
            if (ROWS) s.out_t[frame_in_tile][m] = v;
            else s.out[m][frame_in_tile] = v;
        }
    }
}

// Destination of the operand-row output: the input convolution's A operand of the tensor-core
// Transformer ([rows][80] fp16, time-major) with the chunk bookkeeping of transformer.py:49-64
// folded in — frame g of utterance b lands in every chunk tensor that holds it (at most two:
// chunks overlap by 2 x 50 frames), frame 0 also in the 50 replicate-padding rows of chunk 0.
struct MelRows {
    __half* x0 = nullptr;
    const SeqInfo* seqs = nullptr;   // chunk-major: sequence of (chunk i, utterance b) = i * batch + b
    int batch = 0;
    int chunked = 0, blocks = 1, stride = 0, overlap = 0;
};

template <bool ROWS>
__global__ void __launch_bounds__(kMelThreads, 2)
mel_kernel(const float* __restrict__ audio, int64_t samples, int64_t stride, int frames,
           int tiles_per_row, int total_tiles, MelTables t, __half* __restrict__ mel,
           MelRows rows, int bulk_ok) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MelSmem& s = *reinterpret_cast<MelSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < kNfft; i += kMelThreads) s.window[i] = t.window[i];
    for (int i = tid; i < 7 * 8; i += kMelThreads) {
        const int r = i / 8 + 1, k = i & 7;
        s.tw_pass1[i] = {t.tw512[k * r * 8].x, t.tw512[k * r * 8].y};
    }
    for (int i = tid; i < 7 * 64; i += kMelThreads) {
        const int r = i / 64 + 1, k = i & 63;
        s.tw_pass2[i] = {t.tw512[k * r].x, t.tw512[k * r].y};
    }
    for (int i = tid; i < kBins; i += kMelThreads) s.tw1024[i] = {t.tw1024[i].x, t.tw1024[i].y};
    for (int i = tid; i < t.fb_rounds * kFbPiece * 32; i += kMelThreads) s.fb_w[i] = t.fb_w[i];
    for (int i = tid; i < t.fb_rounds * 32; i += kMelThreads) s.fb_base[i] = t.fb_base[i];
    for (int i = tid; i < kMels; i += kMelThreads) {
        s.band_slot[i] = t.band_slot[i];
        s.band_pieces[i] = t.band_pieces[i];
    }
    if (tid == 0) {
        mbar_init(&s.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t parity = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_row;
        const int f0 = (tile - b * tiles_per_row) * kTileFrames;
        const int nf = min(kTileFrames, frames - f0);
        const int n_samples = (nf - 1) * kHop + kNfft;
        const float* row = audio + (int64_t)b * stride;
        const int64_t first = (int64_t)f0 * kHop - kReflect;   // source index of tile sample 0
        const bool interior = first >= 0 && first + n_samples <= samples;

        if (interior && bulk_ok) {
            // TMA-class bulk copy: one elected thread, completion on the mbarrier.
            if (tid == 0) {
                uint32_t bytes = (uint32_t)n_samples * 4u;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&s.mbar, bytes);
                bulk_copy_g2s(s.audio, row + first, bytes, &s.mbar);
            }
            mbar_wait(&s.mbar, parity);
            parity ^= 1;
        } else {
            for (int i = tid; i < n_samples; i += kMelThreads) {
                int64_t src = first + i;
                if (src < 0) src = -src;                               // reflect (no edge repeat)
                if (src >= samples) src = 2 * (samples - 1) - src;
                s.audio[i] = row[src];
            }
            __syncthreads();
        }

        for (int f = warp; f < nf; f += kMelWarps)
            frame_to_mel<ROWS>(s.audio + f * kHop, s, s.zb[warp], lane, f, t.fb_rounds);
        __syncthreads();

        if (ROWS) {
            // (nf, 80) tile -> operand rows, ten 16-byte segments per frame and destination
            for (int idx = tid; idx < 2 * nf * 10; idx += kMelThreads) {
                const int d = idx / (nf * 10), rem = idx - d * nf * 10;
                const int f = rem / 10, u = rem - f * 10, g = f0 + f;
                int seq = b, local = g;
                if (rows.chunked) {
                    const int chunk = min((g + rows.overlap) / rows.stride, rows.blocks - 1) - d;
                    if (chunk < 0) continue;
                    seq = chunk * rows.batch + b;
                    local = g + rows.overlap - chunk * rows.stride;
                } else if (d) {
                    break;
                }
                const int row0 = rows.seqs[seq].row0, len = rows.seqs[seq].tensor_len;
                if (local < len)
                    *reinterpret_cast<uint4*>(rows.x0 + (int64_t)(row0 + local) * kMels + 8 * u) =
                        *reinterpret_cast<const uint4*>(&s.out_t[f][8 * u]);
            }
            if (rows.chunked && f0 == 0) {   // replicate padding in front of chunk 0 = frame 0
                const int row0 = rows.seqs[b].row0;
                for (int idx = tid; idx < rows.overlap * 10; idx += kMelThreads) {
                    const int local = idx / 10, u = idx - local * 10;
                    *reinterpret_cast<uint4*>(rows.x0 + (int64_t)(row0 + local) * kMels + 8 * u) =
                        *reinterpret_cast<const uint4*>(&s.out_t[0][8 * u]);
                }
            }
            __syncthreads();
            continue;
        }
        // (80, nf) tile -> mel[b][m][f0 + f], contiguous along f
        __half* dst = mel + ((int64_t)b * kMels) * frames + f0;
        if (nf == kTileFrames && (frames & 1) == 0) {
            for (int i = tid; i < kMels * (kTileFrames / 2); i += kMelThreads) {
                int m = i / (kTileFrames / 2), p = i - m * (kTileFrames / 2);
                *reinterpret_cast<__half2*>(dst + (int64_t)m * frames + 2 * p) =
                    *reinterpret_cast<const __half2*>(&s.out[m][2 * p]);
            }
        } else {
            for (int i = tid; i < kMels * kTileFrames; i += kMelThreads) {
                int m = i / kTileFrames, f = i - m * kTileFrames;
                if (f < nf) dst[(int64_t)m * frames + f] = s.out[m][f];
            }
        }
        __syncthreads();   // s.audio / s.out reused by the next tile
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------

// Slaney mel filterbank = librosa.filters.mel(sr=16000, n_fft=1024, n_mels=80)
// (call site ppgs/preprocess/mel.py:61-64): triangles in fp64, stored in fp32,
// scaled by 2/(f[m+2]-f[m]) in fp64 and rounded to fp32 again (librosa keeps the
// weights array in float32).
static void slaney_basis(std::vector<float>& basis) {
    const int n_mels = kMels, n_bins = kBins;
    const double sr = 16000.0, f_sp = 200.0 / 3, min_log_hz = 1000.0;
    const double min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    auto hz_to_mel = [&](double f) {
        return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
    };
    auto mel_to_hz = [&](double m) {
        return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
    };
    std::vector<double> mel_f(n_mels + 2), fft_f(n_bins);
    const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(sr / 2);
    for (int i = 0; i < n_mels + 2; ++i) {
        // numpy.linspace: start + i*step, last point exact
        double m = (i == n_mels + 1) ? m_hi : m_lo + i * ((m_hi - m_lo) / (n_mels + 1));
        mel_f[i] = mel_to_hz(m);
    }
    for (int k = 0; k < n_bins; ++k)
        fft_f[k] = (k == n_bins - 1) ? sr / 2 : k * ((sr / 2) / (n_bins - 1));
    basis.assign((size_t)n_mels * n_bins, 0.f);
    for (int m = 0; m < n_mels; ++m) {
        const double fd0 = mel_f[m + 1] - mel_f[m], fd1 = mel_f[m + 2] - mel_f[m + 1];
        const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
        for (int k = 0; k < n_bins; ++k) {
            double lower = -(mel_f[m] - fft_f[k]) / fd0;
            double upper = (mel_f[m + 2] - fft_f[k]) / fd1;
            double tri = fmax(0.0, fmin(lower, upper));
            float w32 = (float)tri;
            basis[(size_t)m * n_bins + k] = (float)((double)w32 * enorm);
        }
    }
}

int build_mel_tables(ppgs_engine* e, const float* basis_host) {
    std::vector<float> basis;
    if (basis_host) basis.assign(basis_host, basis_host + (size_t)kMels * kBins);
    else slaney_basis(basis);

    std::vector<float> window(kNfft);
    if (!e->host_window.empty()) window = e->host_window;
    else
        for (int n = 0; n < kNfft; ++n)
            window[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * n / kNfft));
    std::vector<float2> tw512(kHalf), tw1024(kBins);
    for (int m = 0; m < kHalf; ++m) {
        double a = -2.0 * M_PI * m / kHalf;
        tw512[m] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int k = 0; k < kBins; ++k) {
        double a = -2.0 * M_PI * k / kNfft;
        tw1024[k] = make_float2((float)cos(a), (float)sin(a));
    }
    std::vector<FilterbankLayout> layout(1);
    if (!build_filterbank_layout(basis.data(), layout.data())) {
        set_error("mel basis needs more than %d pieces of %d bins, more than the kernel holds",
                  kFbMaxRounds * 32, kFbPiece);
        return PPGS_E_INVALID;
    }
    MelTables& t = e->mel;
    auto upload = [&](void** dst, const void* src, size_t bytes) -> int {
        if (*dst) cudaFree(*dst);
        PPGS_CUDA(cudaMalloc(dst, bytes));
        PPGS_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return PPGS_OK;
    };
    PPGS_CHECK(upload((void**)&t.window, window.data(), window.size() * 4));
    PPGS_CHECK(upload((void**)&t.tw512, tw512.data(), tw512.size() * 8));
    PPGS_CHECK(upload((void**)&t.tw1024, tw1024.data(), tw1024.size() * 8));
    const FilterbankLayout& fb = layout[0];
    PPGS_CHECK(upload((void**)&t.fb_w, fb.w, sizeof(fb.w)));
    PPGS_CHECK(upload((void**)&t.fb_base, fb.base, sizeof(fb.base)));
    PPGS_CHECK(upload((void**)&t.band_slot, fb.band_slot, sizeof(fb.band_slot)));
    PPGS_CHECK(upload((void**)&t.band_pieces, fb.band_pieces, sizeof(fb.band_pieces)));
    t.fb_rounds = fb.rounds;
    return PPGS_OK;
}

static int launch_mel_any(ppgs_engine* e, const float* audio, int batch, int64_t samples, int64_t stride,
                          __half* mel, const MelRows* rows, cudaStream_t stream) {
    if (batch <= 0) return PPGS_OK;
    if (samples <= kReflect) {
        // torch reflection_pad1d: "padding size should be less than the input size"
        set_error("mel front-end needs more than %d samples per utterance, got %lld", kReflect,
                  (long long)samples);
        return PPGS_E_INVALID;
    }
    const int frames = (int)(samples / kHop);
    if (frames == 0) return PPGS_OK;
    const int tiles_per_row = (frames + kTileFrames - 1) / kTileFrames;
    const int64_t total = (int64_t)tiles_per_row * batch;
    if (total > INT32_MAX) {
        set_error("too many mel tiles");
        return PPGS_E_TOO_LARGE;
    }
    static PerDeviceOnce attr_set;
    if (attr_set.first(e->device)) {
        PPGS_CUDA(cudaFuncSetAttribute(mel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(MelSmem)));
        PPGS_CUDA(cudaFuncSetAttribute(mel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(MelSmem)));
    }
    // bulk copies need 16-byte aligned sources: base, row stride and tile offset
    // (160*4 and 432*4 are multiples of 16).
    const int bulk_ok = ((reinterpret_cast<uintptr_t>(audio) & 15) == 0) && (stride % 4 == 0);
    const int grid = (int)std::min<int64_t>(total, 2 * (int64_t)e->sm_count);
    {
        LaunchScope scope(e, rows ? "mel_stft_fbank_rows" : "mel_stft_fbank", stream);
        if (rows)
            mel_kernel<true><<<grid, kMelThreads, sizeof(MelSmem), stream>>>(
                audio, samples, stride, frames, tiles_per_row, (int)total, e->mel, nullptr, *rows, bulk_ok);
        else
            mel_kernel<false><<<grid, kMelThreads, sizeof(MelSmem), stream>>>(
                audio, samples, stride, frames, tiles_per_row, (int)total, e->mel, mel, MelRows(), bulk_ok);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

int launch_mel(ppgs_engine* e, const float* audio, int batch, int64_t samples, int64_t stride,
               __half* mel, cudaStream_t stream) {
    return launch_mel_any(e, audio, batch, samples, stride, mel, nullptr, stream);
}

// The mel front-end writing the tensor-core Transformer's input rows directly (no (B, 80, T)
// tensor, no fold pass): `x0` / `seqs_dev` from transformer_tc_input_rows for `plan`.
int launch_mel_rows(ppgs_engine* e, const float* audio, int batch, int64_t samples, int64_t stride,
                    const ForwardPlan& plan, int legacy_mode, __half* x0, const SeqInfo* seqs_dev,
                    cudaStream_t stream) {
    const ppgs_model_config& c = e->cfg;
    if (c.input_channels != kMels || plan.frames != (int)(samples / kHop) || plan.batch != batch) {
        set_error("mel rows: plan does not match the audio batch");
        return PPGS_E_INVALID;
    }
    MelRows rows;
    rows.x0 = x0;
    rows.seqs = seqs_dev;
    rows.batch = batch;
    rows.overlap = c.chunk_overlap;
    rows.stride = c.chunk_length - 2 * c.chunk_overlap;
    rows.chunked = (!legacy_mode && plan.frames > c.chunk_length) ? 1 : 0;
    rows.blocks = rows.chunked ? (plan.frames + rows.stride - 1) / rows.stride : 1;
    // rows behind each chunk tensor (and the pad sequence) are zero operands of the convolution
    PPGS_CUDA(cudaMemsetAsync(x0, 0, (size_t)plan.rows * kMels * sizeof(__half), stream));
    return launch_mel_any(e, audio, batch, samples, stride, nullptr, &rows, stream);
}

}  // namespace ppgs
