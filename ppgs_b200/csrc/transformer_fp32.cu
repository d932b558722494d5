// Transformer forward, PPGS_PRECISION_FP32 mode: every contraction on the CUDA
// cores in fp32 (FFMA).  This is the validation arithmetic of the engine — same
// folded-sequence data layout, masks and epilogues as the tcgen05 path, without
// operand splitting — and the on-device cross-check for the tensor-core kernels.
//
// Replaces ppgs/model/transformer.py:45-81 (+ softmax of ppgs/core.py:593-594).
#include <float.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace ppgs {

// ---------------------------------------------------------------------------
// plan: the chunk bookkeeping of transformer.py:49-64 folded into a batch of
// independent sequences
// ---------------------------------------------------------------------------
int build_plan(const ppgs_engine* e, int batch, int frames, const int64_t* lengths,
               int legacy_mode, ForwardPlan* plan) {
    const ppgs_model_config& c = e->cfg;
    if (batch <= 0 || frames <= 0) {
        set_error("batch and frames must be positive (got %d, %d)", batch, frames);
        return PPGS_E_INVALID;
    }
    int64_t max_len = 0;
    for (int b = 0; b < batch; ++b) {
        if (lengths[b] < 0 || lengths[b] > frames) {
            set_error("lengths[%d]=%lld outside [0, %d]", b, (long long)lengths[b], frames);
            return PPGS_E_INVALID;
        }
        max_len = std::max<int64_t>(max_len, lengths[b]);
    }
    if (max_len != frames) {
        // the reference fails at `input_layer(x) * mask` (transformer.py:75)
        set_error("max(lengths)=%lld must equal the padded length %d", (long long)max_len, frames);
        return PPGS_E_INVALID;
    }
    plan->seqs.clear();
    plan->batch = batch;
    plan->frames = frames;
    plan->max_pitch = 0;
    int row = 0;
    auto add = [&](int b, int tensor_len, int valid, int src_start, int keep_begin, int keep_end,
                   int out_start) {
        SeqInfo s;
        s.row0 = row;
        s.tensor_len = tensor_len;
        s.valid_len = valid;
        s.batch = b;
        s.src_start = src_start;
        s.keep_begin = keep_begin;
        s.keep_end = keep_end;
        s.out_start = out_start;
        const int pitch = (tensor_len + 2 + 127) / 128 * 128;
        plan->max_pitch = std::max(plan->max_pitch, pitch);
        row += pitch;
        plan->seqs.push_back(s);
    };
    const int chunk = c.chunk_length, overlap = c.chunk_overlap, stride = chunk - 2 * overlap;
    if (legacy_mode || frames <= chunk) {
        if (frames > c.max_len) {
            set_error("size is too large");   // PositionalEncoding.forward, transformer.py:103-104
            return PPGS_E_TOO_LARGE;
        }
        for (int b = 0; b < batch; ++b) add(b, frames, (int)lengths[b], 0, 0, frames, 0);
    } else {
        const int blocks = (frames + stride - 1) / stride;
        std::vector<int64_t> rem(lengths, lengths + batch);
        for (int i = 0; i < blocks; ++i) {
            const int start = i * stride;                                  // in left-padded coordinates
            const int stop = std::min((i + 1) * stride + 2 * overlap, frames + overlap);
            const int tensor_len = stop - start;
            for (int b = 0; b < batch; ++b) {
                int64_t cl = std::min<int64_t>(std::max<int64_t>(rem[b] + overlap, 0), chunk);
                if (cl == overlap) cl = 0;
                rem[b] = std::max<int64_t>(rem[b] - stride, 0);
                const int keep_end = std::min(chunk - overlap, tensor_len);
                add(b, tensor_len, (int)std::min<int64_t>(cl, tensor_len), start - overlap, overlap,
                    std::max(keep_end, overlap), start);
            }
        }
    }
    if (row % 256) {
        // the CTA-pair GEMMs work on pairs of 128-row tiles: pad with one empty sequence
        add(0, 0, 0, 0, 0, 0, 0);
    }
    plan->rows = row;
    return PPGS_OK;
}

// ---------------------------------------------------------------------------
// fold: (B, C, T) fp16 features -> time-major rows [rows][C] fp32, replicate left
// padding (transformer.py:54), zero rows beyond each chunk tensor
// ---------------------------------------------------------------------------
__global__ void fold_kernel(const __half* __restrict__ feats, int C, int T,
                            const SeqInfo* __restrict__ seqs, const int* __restrict__ tile_seq,
                            float* __restrict__ x0) {
    __shared__ float tile[32][33];
    const int row_base = blockIdx.x * 32, c_base = blockIdx.y * 32;
    const SeqInfo s = seqs[tile_seq[row_base >> 7]];
    const int tx = threadIdx.x, ty = threadIdx.y;   // (32, 8)
    for (int i = ty; i < 32; i += 8) {
        const int c = c_base + i, t = row_base - s.row0 + tx;
        float v = 0.f;
        if (c < C && t < s.tensor_len) {
            int f = max(s.src_start + t, 0);
            v = __half2float(feats[((int64_t)s.batch * C + c) * T + f]);
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c_base + tx;
        if (c < C) x0[(int64_t)(row_base + i) * C + c] = tile[tx][i];
    }
}

// ---------------------------------------------------------------------------
// SGEMM: out[M,N] = A[M,K] (row stride lda, rows may overlap: conv-as-GEMM) * B[N,K]^T
// ---------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256) sgemm_kernel(SgemmArgs a) {
    constexpr int BM = 128, BN = 64, BK = 16;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int tm = (tid >> 4) * 8, tn = (tid & 15) * 4;   // 16x16 threads, 8x4 outputs each
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < a.K; k0 += BK) {
        // A tile: 128 rows x 16 k = 512 float4, 2 per thread
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int idx = tid + i * 256;
            int r = idx >> 2, kq = (idx & 3) * 4;
            float4 v = *reinterpret_cast<const float4*>(a.A + (int64_t)(m0 + r) * a.lda + k0 + kq);
            As[kq + 0][r] = v.x;
            As[kq + 1][r] = v.y;
            As[kq + 2][r] = v.z;
            As[kq + 3][r] = v.w;
        }
        {
            int r = tid >> 2, kq = (tid & 3) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + r < a.N) v = *reinterpret_cast<const float4*>(a.B + (int64_t)(n0 + r) * a.K + k0 + kq);
            Bs[kq + 0][r] = v.x;
            Bs[kq + 1][r] = v.y;
            Bs[kq + 2][r] = v.z;
            Bs[kq + 3][r] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[k][tm]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[k][tm + 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    SeqInfo s;
    if (EPI == EPI_IN) s = a.seqs[a.tile_seq[m0 >> 7]];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + tm + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn + j;
            if (n >= a.N) continue;
            float v = acc[i][j] + a.bias[n];
            if (EPI == EPI_BIAS_RELU) v = fmaxf(v, 0.f);
            if (EPI == EPI_BIAS_GELU) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
            if (EPI == EPI_BIAS_RES) v += a.res[(int64_t)m * a.N + n];
            if (EPI == EPI_IN) {
                const int t = m - s.row0;
                if (t >= s.tensor_len) v = 0.f;
                else v = (t < s.valid_len ? v : 0.f) + a.pe[(int64_t)t * a.N + n];
            }
            a.out[(int64_t)m * a.ldo + n] = v;
        }
    }
}

template <int EPI>
static int launch_sgemm(ppgs_engine* e, const char* name, const SgemmArgs& a,
                        cudaStream_t stream) {
    dim3 grid(a.M / 128, (a.N + 63) / 64);
    {
        LaunchScope scope(e, name, stream);
        sgemm_kernel<EPI><<<grid, 256, 0, stream>>>(a);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

int launch_sgemm_any(ppgs_engine* e, const char* name, int epi, const SgemmArgs& a,
                     cudaStream_t stream) {
    if (a.M % 128 || a.K % 16 || a.lda % 4) {
        set_error("sgemm %s: M %% 128, K %% 16 and lda %% 4 must be 0 (M=%d K=%d)", name, a.M, a.K);
        return PPGS_E_INVALID;
    }
    switch (epi) {
        case EPI_BIAS: return launch_sgemm<EPI_BIAS>(e, name, a, stream);
        case EPI_BIAS_RELU: return launch_sgemm<EPI_BIAS_RELU>(e, name, a, stream);
        case EPI_BIAS_RES: return launch_sgemm<EPI_BIAS_RES>(e, name, a, stream);
        case EPI_BIAS_GELU: return launch_sgemm<EPI_BIAS_GELU>(e, name, a, stream);
        default: set_error("sgemm: unknown epilogue %d", epi); return PPGS_E_INVALID;
    }
}

// ---------------------------------------------------------------------------
// LayerNorm over the hidden dimension, one warp per row; padding rows -> 0
// ---------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ y, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, const SeqInfo* __restrict__ seqs,
                 const int* __restrict__ tile_seq, int rows, float* __restrict__ x) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const SeqInfo s = seqs[tile_seq[row >> 7]];
    constexpr int PER = H / 32;
    float v[PER];
    float* dst = x + (int64_t)row * H;
    if (row - s.row0 >= s.tensor_len) {
#pragma unroll
        for (int i = 0; i < PER; ++i) dst[lane + 32 * i] = 0.f;
        return;
    }
    const float* src = y + (int64_t)row * H;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        v[i] = src[lane + 32 * i];
        sum += v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / H;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        float d = v[i] - mean;
        sq = fmaf(d, d, sq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / H + eps);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        int c = lane + 32 * i;
        dst[c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
    }
}

// ---------------------------------------------------------------------------
// attention, fp32: block = (32 queries, one head, one sequence); warp = 4 queries;
// lane = key within a 32-key tile for the scores, = feature slice for P.V
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
attention_fp32_kernel(const float* __restrict__ qkv, int H, const SeqInfo* __restrict__ seqs,
                      int causal, float scale, float* __restrict__ out) {
    constexpr int QT = 32, KT = 32, DP = D + 1, PER = D / 32;
    extern __shared__ float smem[];
    float* qs = smem;                // [QT][D]
    float* ks = qs + QT * D;         // [KT][DP]
    float* vs = ks + KT * DP;        // [KT][D]
    const SeqInfo s = seqs[blockIdx.z];
    const int q0 = blockIdx.x * QT;
    if (q0 >= s.tensor_len) return;
    const int head = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t ld = 3 * H;
    const float* base = qkv + (int64_t)s.row0 * ld;

    for (int i = tid; i < QT * D; i += 256) {
        int q = i / D, d = i - q * D;
        qs[i] = base[(int64_t)(q0 + q) * ld + head * D + d];   // rows < pitch always allocated
    }
    float m[4], l[4], acc[4][PER];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        m[q] = -FLT_MAX;
        l[q] = 0.f;
#pragma unroll
        for (int c = 0; c < PER; ++c) acc[q][c] = 0.f;
    }
    int kend = s.valid_len;
    if (causal) kend = min(kend, q0 + QT);
    for (int k0 = 0; k0 < kend; k0 += KT) {
        __syncthreads();
        for (int i = tid; i < KT * D; i += 256) {
            int j = i / D, d = i - j * D;
            const float* r = base + (int64_t)(k0 + j) * ld + head * D + d;
            ks[j * DP + d] = r[H];
            vs[j * D + d] = r[2 * H];
        }
        __syncthreads();
        const int key = k0 + lane;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int t = q0 + warp * 4 + q;
            const float* qr = qs + (warp * 4 + q) * D;
            float sc = 0.f;
#pragma unroll 8
            for (int d = 0; d < D; ++d) sc = fmaf(qr[d], ks[lane * DP + d], sc);
            sc *= scale;
            const bool ok = key < s.valid_len && (!causal || key <= t);
            sc = ok ? sc : -FLT_MAX;
            float tmax = sc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
            const float mnew = fmaxf(m[q], tmax);
            const float p = ok ? expf(sc - mnew) : 0.f;
            const float corr = expf(m[q] - mnew);   // m == -FLT_MAX, mnew finite -> 0
            float psum = p;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
            l[q] = l[q] * corr + psum;
            m[q] = mnew;
#pragma unroll
            for (int c = 0; c < PER; ++c) acc[q][c] *= corr;
            for (int j = 0; j < KT; ++j) {
                const float pj = __shfl_sync(0xffffffffu, p, j);
#pragma unroll
                for (int c = 0; c < PER; ++c) acc[q][c] = fmaf(pj, vs[j * D + lane + 32 * c], acc[q][c]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int t = q0 + warp * 4 + q;
        const float inv = l[q] > 0.f ? 1.f / l[q] : 0.f;   // fully masked row -> zeros
        float* dst = out + (int64_t)(s.row0 + t) * H + head * D;
#pragma unroll
        for (int c = 0; c < PER; ++c) dst[lane + 32 * c] = acc[q][c] * inv;
    }
}

// ---------------------------------------------------------------------------
// finalize: mask (transformer.py:81), softmax over channels (core.py:593-594),
// un-chunk (`[..., 50:450]` + cat, transformer.py:63-64) into (B, O, T)
// ---------------------------------------------------------------------------
__global__ void finalize_kernel(const float* __restrict__ logits, int ldl, int O,
                                const SeqInfo* __restrict__ seqs, int T, int softmax,
                                float* __restrict__ out) {
    const SeqInfo s = seqs[blockIdx.y];
    const int t = s.keep_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= s.keep_end) return;
    const float* src = logits + (int64_t)(s.row0 + t) * ldl;
    const bool valid = t < s.valid_len;
    float* dst = out + (int64_t)s.batch * O * T + s.out_start + (t - s.keep_begin);
    if (!softmax) {
        for (int c = 0; c < O; ++c) dst[(int64_t)c * T] = valid ? src[c] : 0.f;
        return;
    }
    float mx = -FLT_MAX;
    for (int c = 0; c < O; ++c) mx = fmaxf(mx, valid ? src[c] : 0.f);
    float sum = 0.f;
    for (int c = 0; c < O; ++c) sum += expf((valid ? src[c] : 0.f) - mx);
    const float inv = 1.f / sum;
    for (int c = 0; c < O; ++c) dst[(int64_t)c * T] = expf((valid ? src[c] : 0.f) - mx) * inv;
}

int launch_finalize(ppgs_engine* e, const float* logits, int ldl, const ForwardPlan& plan,
                    const SeqInfo* seqs_dev, int softmax, float* out, cudaStream_t stream) {
    int max_keep = 0;
    for (const SeqInfo& s : plan.seqs) max_keep = std::max(max_keep, s.keep_end - s.keep_begin);
    if (max_keep == 0) return PPGS_OK;
    dim3 grid((max_keep + 127) / 128, (unsigned)plan.seqs.size());
    {
        LaunchScope scope(e, "finalize_softmax_unchunk", stream);
        finalize_kernel<<<grid, 128, 0, stream>>>(logits, ldl, e->cfg.output_channels, seqs_dev,
                                                  plan.frames, softmax, out);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

int launch_fold(ppgs_engine* e, const __half* feats, const ForwardPlan& plan,
                const SeqInfo* seqs_dev, const int* tile_seq_dev, float* x0, cudaStream_t stream) {
    const int C = e->cfg.input_channels;
    dim3 grid(plan.rows / 32, (C + 31) / 32);
    {
        LaunchScope scope(e, "fold_chunks", stream);
        fold_kernel<<<grid, dim3(32, 8), 0, stream>>>(feats, C, plan.frames, seqs_dev,
                                                      tile_seq_dev, x0);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

template <int H>
static int launch_ln(ppgs_engine* e, const float* y, const float* g, const float* b,
                     const SeqInfo* seqs, const int* tile_seq, int rows, float* x,
                     cudaStream_t stream) {
    {
        LaunchScope scope(e, "layernorm_fp32", stream);
        layernorm_kernel<H><<<(rows + 7) / 8, 256, 0, stream>>>(
            y, g, b, e->cfg.layer_norm_eps, seqs, tile_seq, rows, x);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

template <int D>
static int launch_attention(ppgs_engine* e, const float* qkv, const ForwardPlan& plan,
                            const SeqInfo* seqs, float* out, cudaStream_t stream) {
    const int H = e->cfg.hidden_channels;
    const size_t smem = (size_t)(32 * D + 32 * (D + 1) + 32 * D) * sizeof(float);
    static PerDeviceOnce attr;
    if (attr.first(e->device)) {
        PPGS_CUDA(cudaFuncSetAttribute(attention_fp32_kernel<D>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid(plan.max_pitch / 32, e->cfg.num_heads, (unsigned)plan.seqs.size());
    {
        LaunchScope scope(e, "attention_fp32", stream);
        attention_fp32_kernel<D><<<grid, 256, smem, stream>>>(qkv, H, seqs, e->cfg.is_causal,
                                                              1.f / sqrtf((float)D), out);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

int launch_attention_fp32_any(ppgs_engine* e, int head_dim, const float* qkv, int H, int heads,
                              int max_pitch, int nseq, const SeqInfo* seqs, int causal,
                              float* out, cudaStream_t stream) {
    auto run = [&](auto kernel, int D) -> int {
        const size_t smem = (size_t)(32 * D + 32 * (D + 1) + 32 * D) * sizeof(float);
        PPGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(max_pitch / 32, heads, (unsigned)nseq);
        {
            LaunchScope scope(e, "attention_fp32", stream);
            kernel<<<grid, 256, smem, stream>>>(qkv, H, seqs, causal, 1.f / sqrtf((float)D), out);
        }
        PPGS_CUDA(cudaGetLastError());
        return PPGS_OK;
    };
    if (head_dim == 64) return run(attention_fp32_kernel<64>, 64);
    if (head_dim == 128) return run(attention_fp32_kernel<128>, 128);
    if (head_dim == 256) return run(attention_fp32_kernel<256>, 256);
    set_error("attention_fp32: head_dim %d not built", head_dim);
    return PPGS_E_UNSUPPORTED;
}

int transformer_forward_fp32(ppgs_engine* e, const __half* features, const ForwardPlan& plan,
                             int softmax, float* out, cudaStream_t stream) {
    const ppgs_model_config& c = e->cfg;
    const int C = c.input_channels, H = c.hidden_channels, F = c.ffn_channels;
    const int O = c.output_channels, k = c.kernel_size, half = k / 2;
    const int rows = plan.rows, D = H / c.num_heads;
    if (!((H == 256 || H == 512) && (D == 128 || D == 256))) {
        set_error("fp32 path supports hidden 256/512 with head_dim 128/256 (got %d, %d)", H, D);
        return PPGS_E_UNSUPPORTED;
    }
    if (C % 4 || H % 4 || F % 16 || (k * C) % 16 || (k * H) % 16) {
        set_error("channel counts must keep K a multiple of 16");
        return PPGS_E_UNSUPPORTED;
    }
    // workspace carve-up (floats); `guard` zero rows before and after x0 / x for the
    // 'same' convolution halo
    const size_t guard = half;
    Carver w;
    const size_t o_x0 = w.take((rows + 2 * guard) * (size_t)C * 4);
    const size_t o_x = w.take((rows + 2 * guard) * (size_t)H * 4);
    const size_t o_y = w.take((size_t)rows * H * 4);
    const size_t o_qkv = w.take((size_t)rows * 3 * H * 4);
    const size_t o_att = w.take((size_t)rows * H * 4);
    const size_t o_ff = w.take((size_t)rows * F * 4);
    const size_t o_logits = w.take((size_t)rows * O * 4);
    const size_t o_seqs = w.take(plan.seqs.size() * sizeof(SeqInfo));
    const size_t o_tiles = w.take((size_t)(rows / 128) * 4);
    PPGS_CHECK(ensure_workspace(e, w.off));
    char* ws = static_cast<char*>(e->workspace);
    float* x0 = reinterpret_cast<float*>(ws + o_x0);
    float* x = reinterpret_cast<float*>(ws + o_x);
    float* y = reinterpret_cast<float*>(ws + o_y);
    float* qkv = reinterpret_cast<float*>(ws + o_qkv);
    float* att = reinterpret_cast<float*>(ws + o_att);
    float* ff = reinterpret_cast<float*>(ws + o_ff);
    float* logits = reinterpret_cast<float*>(ws + o_logits);
    SeqInfo* seqs_dev = reinterpret_cast<SeqInfo*>(ws + o_seqs);
    int* tile_seq_dev = reinterpret_cast<int*>(ws + o_tiles);
    PPGS_CHECK(upload_plan(e, plan, seqs_dev, tile_seq_dev, stream));

    PPGS_CUDA(cudaMemsetAsync(x0, 0, guard * C * sizeof(float), stream));
    PPGS_CUDA(cudaMemsetAsync(x0 + (rows + guard) * (size_t)C, 0, guard * C * sizeof(float), stream));
    PPGS_CUDA(cudaMemsetAsync(x, 0, guard * H * sizeof(float), stream));
    PPGS_CUDA(cudaMemsetAsync(x + (rows + guard) * (size_t)H, 0, guard * H * sizeof(float), stream));
    float* x0r = x0 + guard * C;   // row 0
    float* xr = x + guard * H;

    PPGS_CHECK(launch_fold(e, features, plan, seqs_dev, tile_seq_dev, x0r, stream));

    SgemmArgs a{};
    a.seqs = seqs_dev;
    a.tile_seq = tile_seq_dev;
    // input conv: im2col rows are contiguous 5*C floats starting `half` rows earlier
    a.A = x0r - (int64_t)half * C; a.lda = C; a.B = e->conv_in_w; a.bias = e->conv_in_b;
    a.out = xr; a.ldo = H; a.M = rows; a.N = H; a.K = k * C; a.pe = e->pe;
    PPGS_CHECK(launch_sgemm<EPI_IN>(e, "sgemm_conv_in", a, stream));

    for (int layer = 0; layer < c.num_layers; ++layer) {
        const LayerWeights& L = e->layers[layer];
        a.A = xr; a.lda = H; a.B = L.in_w; a.bias = L.in_b; a.out = qkv; a.ldo = 3 * H;
        a.N = 3 * H; a.K = H;
        PPGS_CHECK(launch_sgemm<EPI_BIAS>(e, "sgemm_qkv", a, stream));
        if (D == 128) PPGS_CHECK(launch_attention<128>(e, qkv, plan, seqs_dev, att, stream));
        else PPGS_CHECK(launch_attention<256>(e, qkv, plan, seqs_dev, att, stream));
        a.A = att; a.lda = H; a.B = L.out_w; a.bias = L.out_b; a.out = y; a.ldo = H; a.N = H;
        a.K = H; a.res = xr;
        PPGS_CHECK(launch_sgemm<EPI_BIAS_RES>(e, "sgemm_out_proj", a, stream));
        if (H == 256) PPGS_CHECK(launch_ln<256>(e, y, L.n1_w, L.n1_b, seqs_dev, tile_seq_dev, rows, xr, stream));
        else PPGS_CHECK(launch_ln<512>(e, y, L.n1_w, L.n1_b, seqs_dev, tile_seq_dev, rows, xr, stream));
        a.A = xr; a.lda = H; a.B = L.l1_w; a.bias = L.l1_b; a.out = ff; a.ldo = F; a.N = F; a.K = H;
        PPGS_CHECK(launch_sgemm<EPI_BIAS_RELU>(e, "sgemm_ffn1", a, stream));
        a.A = ff; a.lda = F; a.B = L.l2_w; a.bias = L.l2_b; a.out = y; a.ldo = H; a.N = H; a.K = F;
        a.res = xr;
        PPGS_CHECK(launch_sgemm<EPI_BIAS_RES>(e, "sgemm_ffn2", a, stream));
        if (H == 256) PPGS_CHECK(launch_ln<256>(e, y, L.n2_w, L.n2_b, seqs_dev, tile_seq_dev, rows, xr, stream));
        else PPGS_CHECK(launch_ln<512>(e, y, L.n2_w, L.n2_b, seqs_dev, tile_seq_dev, rows, xr, stream));
    }
    a.A = xr - (int64_t)half * H; a.lda = H; a.B = e->conv_out_w; a.bias = e->conv_out_b;
    a.out = logits; a.ldo = O; a.N = O; a.K = k * H;
    PPGS_CHECK(launch_sgemm<EPI_BIAS>(e, "sgemm_conv_out", a, stream));
    PPGS_CHECK(launch_finalize(e, logits, O, plan, seqs_dev, softmax, out, stream));
    return PPGS_OK;
}

}  // namespace ppgs
