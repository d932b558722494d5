// `w2v2fb` representation: wav2vec2-base latents, nearest-upsampled to the PPG frame
// rate — replaces ppgs.preprocess.w2v2fb.from_audios (ppgs/preprocess/w2v2fb/core.py:32-75)
// and the Hugging Face `Wav2Vec2Model` forward it calls (transformers, un-vendored:
// modeling_wav2vec2.py feature encoder :254-324,:382-419, projection :422-436, positional
// conv :326-380, encoder layers :576-610, encoder :658-728, mask reduction :1005-1044).
//
// The convolutional feature encoder, the positional convolution and the attention run in
// the CUDA-core fp32 arithmetic (SGEMM / attention kernels of transformer_fp32.cu); the
// encoder projections and FFN go through the split-fp16 tcgen05 GEMM (gemm_tc.cu).  All
// activations are time-major [rows][C]; utterance b owns rows [b*P_l, b*P_l + T_l) of
// layer l with pitches P_{l-1} = 2 P_l, so a stride-2 convolution is a GEMM whose A rows
// start every 2*C floats and overlap (lda = stride*C, K = kernel*C).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <string>

#include "attention_tc.cuh"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace ppgs {

namespace {

constexpr int kConvDim = 512, kHidden = 768, kHeads = 12, kLayers = 12, kFfn = 3072;
constexpr int kPosKernel = 128, kPosGroups = 16, kPosPer = kHidden / kPosGroups;   // 48
constexpr int kNumConv = 7;
const int kConvKernel[kNumConv] = {10, 3, 3, 3, 3, 2, 2};
const int kConvStride[kNumConv] = {5, 2, 2, 2, 2, 2, 2};
constexpr int kW2v2Pad = 40;   // w2v2fb/core.py:54

// ---- layer 0: Conv1d(1, 512, k=10, s=5), no bias.  Block = 32 frames x 512 channels.
__global__ void __launch_bounds__(256)
conv0_kernel(const float* __restrict__ audio, int64_t stride, int samples, int T0, int64_t P0,
             const float* __restrict__ w /* [512][10] */, float* __restrict__ out /* [B*P0][512] */) {
    __shared__ float xs[32 * 5 + 5];
    __shared__ float ws[kConvDim * 10];
    const int b = blockIdx.y, t0 = blockIdx.x * 32, tid = threadIdx.x;
    for (int i = tid; i < kConvDim * 10; i += 256) ws[i] = w[i];
    for (int i = tid; i < 32 * 5 + 5; i += 256) {
        const int64_t pos = (int64_t)t0 * 5 + i - kW2v2Pad;   // index into the un-padded audio
        xs[i] = (pos >= 0 && pos < samples) ? audio[(int64_t)b * stride + pos] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < 32 * kConvDim; i += 256) {
        const int tt = i / kConvDim, c = i - tt * kConvDim;
        if (t0 + tt >= T0) continue;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 10; ++j) acc = fmaf(ws[c * 10 + j], xs[tt * 5 + j], acc);
        out[((int64_t)b * P0 + t0 + tt) * kConvDim + c] = acc;
    }
}

// ---- GroupNorm(512 groups of 1 channel) statistics over time: double sums per (b, c)
__global__ void __launch_bounds__(512)
groupnorm_stats_kernel(const float* __restrict__ x, int T0, int64_t P0, double* __restrict__ sums /* [B][512][2] */) {
    const int b = blockIdx.y, c = threadIdx.x;
    const int t_begin = blockIdx.x * 256, t_end = min(t_begin + 256, T0);
    float s = 0.f, q = 0.f;
    for (int t = t_begin; t < t_end; ++t) {
        const float v = x[((int64_t)b * P0 + t) * kConvDim + c];
        s += v;
        q = fmaf(v, v, q);
    }
    atomicAdd(&sums[((int64_t)b * kConvDim + c) * 2], (double)s);
    atomicAdd(&sums[((int64_t)b * kConvDim + c) * 2 + 1], (double)q);
}

__device__ __forceinline__ float gelu_exact(float v) {
    return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
}

__global__ void __launch_bounds__(512)
groupnorm_gelu_kernel(float* __restrict__ x, int T0, int64_t P0, const double* __restrict__ sums,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
    const int b = blockIdx.y, c = threadIdx.x;
    const double mean = sums[((int64_t)b * kConvDim + c) * 2] / T0;
    const double var = sums[((int64_t)b * kConvDim + c) * 2 + 1] / T0 - mean * mean;
    const float m = (float)mean, r = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma[c], be = beta[c];
    const int t_begin = blockIdx.x * 64, t_end = min(t_begin + 64, T0);
    for (int t = t_begin; t < t_end; ++t) {
        float* p = x + ((int64_t)b * P0 + t) * kConvDim + c;
        *p = gelu_exact((*p - m) * r * g + be);
    }
}

// ---- tensor-core path: conv0 is 10 MACs per output, so the 2.1 GB fp32 activation of a
// 32 x 10 s batch is never materialised: the statistics pass and the normalise + GELU +
// split pass both recompute it from the audio (same fmaf chain as conv0_kernel, same
// accumulation order as groupnorm_stats_kernel -> same values), thread = channel.
__device__ __forceinline__ void conv0_stage(const float* __restrict__ audio, int64_t stride, int samples, int b,
                                            int t0, int frames, float* xs) {
    for (int i = threadIdx.x; i < frames * 5 + 5; i += blockDim.x) {
        const int64_t pos = (int64_t)t0 * 5 + i - kW2v2Pad;
        xs[i] = (pos >= 0 && pos < samples) ? audio[(int64_t)b * stride + pos] : 0.f;
    }
}

__global__ void __launch_bounds__(512)
conv0_stats_kernel(const float* __restrict__ audio, int64_t stride, int samples, int T0,
                   const float* __restrict__ w /* [512][10] */, double* __restrict__ sums) {
    __shared__ float xs[256 * 5 + 5];
    const int b = blockIdx.y, c = threadIdx.x, t_begin = blockIdx.x * 256;
    const int frames = min(256, T0 - t_begin);
    conv0_stage(audio, stride, samples, b, t_begin, frames, xs);
    float wc[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) wc[j] = w[c * 10 + j];
    __syncthreads();
    float s = 0.f, q = 0.f;
    for (int tt = 0; tt < frames; ++tt) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 10; ++j) v = fmaf(wc[j], xs[tt * 5 + j], v);
        s += v;
        q = fmaf(v, v, q);
    }
    atomicAdd(&sums[((int64_t)b * kConvDim + c) * 2], (double)s);
    atomicAdd(&sums[((int64_t)b * kConvDim + c) * 2 + 1], (double)q);
}

// thread = channel pair (half2 stores: 128 contiguous bytes per warp, row and plane)
__global__ void __launch_bounds__(256)
conv0_groupnorm_gelu_planes_kernel(const float* __restrict__ audio, int64_t stride, int samples, int T0,
                                   int64_t P0, int64_t plane_stride, const float* __restrict__ w,
                                   const double* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, __half* __restrict__ planes) {
    __shared__ float xs[64 * 5 + 5];
    const int b = blockIdx.y, c = 2 * threadIdx.x, t_begin = blockIdx.x * 64;
    const int t_end = (int)min((int64_t)t_begin + 64, P0);
    conv0_stage(audio, stride, samples, b, t_begin, 64, xs);
    float wc[2][10], m[2], r[2], g[2], be[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int j = 0; j < 10; ++j) wc[h][j] = w[(c + h) * 10 + j];
        const double mean = sums[((int64_t)b * kConvDim + c + h) * 2] / T0;
        const double var = sums[((int64_t)b * kConvDim + c + h) * 2 + 1] / T0 - mean * mean;
        m[h] = (float)mean;
        r[h] = (float)(1.0 / sqrt(var + (double)eps));
        g[h] = gamma[c + h];
        be[h] = beta[c + h];
    }
    __syncthreads();
#pragma unroll 2
    for (int t = t_begin; t < t_end; ++t) {   // rows [T0, P0) -> 0: finite inputs for the junk rows
        const int64_t at = ((int64_t)b * P0 + t) * kConvDim + c;
        uint32_t hi2 = 0u, lo2 = 0u;
        if (t < T0) {
            float y[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float v = 0.f;
#pragma unroll
                for (int j = 0; j < 10; ++j) v = fmaf(wc[h][j], xs[(t - t_begin) * 5 + j], v);
                y[h] = gelu_exact((v - m[h]) * r[h] * g[h] + be[h]);
            }
            tc::split2_f16(y[0], y[1], hi2, lo2);
        }
        *reinterpret_cast<uint32_t*>(planes + at) = hi2;
        *reinterpret_cast<uint32_t*>(planes + plane_stride + at) = lo2;
    }
}

// ---- feature-projection LayerNorm(512): conv rows (pitch Q, split planes or fp32) -> fp32
// rows at the encoder pitch P; rows t >= T of a sequence and the dummy tail rows -> 0
template <bool PLANES>
__global__ void __launch_bounds__(256)
feature_layernorm_kernel(const void* __restrict__ src, int64_t plane_stride, int64_t Q, int64_t P,
                         int T, int batch, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, int rows, float* __restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    constexpr int H = kConvDim, PER = H / 32;
    const int b = (int)(row / P), t = (int)(row - (int64_t)b * P);
    float* dst = out + (int64_t)row * H;
    if (b >= batch || t >= T) {
#pragma unroll
        for (int i = 0; i < PER; ++i) dst[lane + 32 * i] = 0.f;
        return;
    }
    const int64_t base = ((int64_t)b * Q + t) * H;
    float v[PER];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int64_t at = base + lane + 32 * i;
        if (PLANES) {
            const __half* x = static_cast<const __half*>(src);
            v[i] = __half2float(x[at]) + __half2float(x[plane_stride + at]);
        } else {
            v[i] = static_cast<const float*>(src)[at];
        }
        sum += v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / H;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const float d = v[i] - mean;
        sq = fmaf(d, d, sq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / H + eps);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = lane + 32 * i;
        dst[c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
    }
}

// ---- out = LayerNorm(a [+ b]) per row; rows with t >= limit[b] (optional) -> 0
template <int H>
__global__ void __launch_bounds__(256)
add_layernorm_kernel(const float* a /* may alias out */, const float* __restrict__ b2,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     int rows, float* out, int64_t pitch_a = 0, int64_t pitch_b = 0,
                     int batch = 0) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    constexpr int PER = H / 32;
    float v[PER];
    float sum = 0.f;
    // b2 rows of sequence s start at s * pitch_b instead of s * pitch_a (positional conv output)
    int64_t row_b = row;
    bool has_b = b2 != nullptr;
    if (pitch_b) {
        const int64_t s = row / pitch_a;
        row_b = s * pitch_b + (row - s * pitch_a);
        has_b = has_b && s < batch;
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int64_t at = (int64_t)row * H + lane + 32 * i;
        v[i] = a[at] + (has_b ? b2[row_b * H + lane + 32 * i] : 0.f);
        sum += v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / H;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const float d = v[i] - mean;
        sq = fmaf(d, d, sq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / H + eps);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = lane + 32 * i;
        out[(int64_t)row * H + c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
    }
}

// ---- zero the rows that are padding tokens or beyond the sequence (modeling :681-684)
__global__ void mask_rows_kernel(float* __restrict__ h, int64_t P, const int* __restrict__ out_len,
                                 int C, int batch) {
    const int64_t row = blockIdx.x;
    const int b = (int)(row / P), t = (int)(row - (int64_t)b * P);
    if (b < batch && t < out_len[b]) return;
    for (int c = threadIdx.x; c < C; c += blockDim.x) h[row * C + c] = 0.f;
}

// ---- h (pitch P) -> group-major copy Xg [16][64 + Mg + 64][48] at pitch Pg = P + 128: the
// zero rows between utterances are the halo of the k=128 positional convolution
__global__ void group_major_kernel(const float* __restrict__ h, int64_t P, int64_t Pg, int64_t Mg,
                                   float* __restrict__ xg) {
    const int64_t row = blockIdx.x;
    const int64_t b = row / P, dst = 64 + b * Pg + (row - b * P);
    for (int c = threadIdx.x; c < kHidden; c += blockDim.x) {
        const int g = c / kPosPer, i = c - g * kPosPer;
        xg[((int64_t)g * (Mg + 128) + dst) * kPosPer + i] = h[row * kHidden + c];
    }
}

// same copy as split-fp16 planes [2][16 * (Mg + 128)][48] (tensor-core positional conv)
__global__ void group_major_planes_kernel(const float* __restrict__ h, int64_t P, int64_t Pg, int64_t Mg,
                                          __half* __restrict__ xg) {
    const int64_t row = blockIdx.x;
    const int64_t b = row / P, dst = 64 + b * Pg + (row - b * P);
    const int64_t plane_stride = (int64_t)kPosGroups * (Mg + 128) * kPosPer;
    for (int c = threadIdx.x; c < kHidden; c += blockDim.x) {
        const int g = c / kPosPer, i = c - g * kPosPer;
        __half hi, lo;
        tc::split_f16(h[row * kHidden + c], hi, lo);
        const int64_t at = ((int64_t)g * (Mg + 128) + dst) * kPosPer + i;
        xg[at] = hi;
        xg[plane_stride + at] = lo;
    }
}

// ---- (B*P rows, 768) hidden -> (B, 768, frames) fp16, nearest neighbour in time
__global__ void upsample_kernel(const float* __restrict__ h, int64_t P, int T6, int frames,
                                float scale, __half* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, f0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // (32, 8)
    for (int i = ty; i < 32; i += 8) {
        const int f = f0 + i;
        float v = 0.f;
        if (f < frames) {
            const int src = min((int)floorf((float)f * scale), T6 - 1);
            v = h[((int64_t)b * P + src) * kHidden + c0 + tx];
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int f = f0 + tx, c = c0 + i;
        if (f < frames) out[((int64_t)b * kHidden + c) * frames + f] = __float2half_rn(tile[tx][i]);
    }
}

// ---- fp32 rows -> split-fp16 planes [2][M][C]
__global__ void to_planes_kernel(const float* __restrict__ x, int64_t count, __half* __restrict__ planes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    __half hi, lo;
    tc::split_f16(x[i], hi, lo);
    planes[i] = hi;
    planes[count + i] = lo;
}

// ---- x <- LayerNorm(x + y) over split planes (one warp per row), optional fp32 copy
template <int H>
__global__ void __launch_bounds__(256)
add_layernorm_planes_kernel(__half* __restrict__ x, const __half* __restrict__ y, const __half* __restrict__ y2,
                            int64_t plane_stride, const float* __restrict__ gamma,
                            const float* __restrict__ beta, float eps, int rows, float* __restrict__ out_f32) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    constexpr int PER = H / 64;   // half2 per lane
    float v[2 * PER];
    float sum = 0.f;
    const int64_t base = (int64_t)row * H;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int64_t at = base + 2 * (lane + 32 * i);
        const float2 xh = __half22float2(*reinterpret_cast<const __half2*>(x + at));
        const float2 xl = __half22float2(*reinterpret_cast<const __half2*>(x + plane_stride + at));
        const float2 yh = __half22float2(*reinterpret_cast<const __half2*>(y + at));
        const float2 yl = __half22float2(*reinterpret_cast<const __half2*>(y + plane_stride + at));
        float y0 = yh.x + yl.x, y1 = yh.y + yl.y;
        if (y2) {   // second partial sum of a K-split projection (fp32 add: round to nearest)
            const float2 zh = __half22float2(*reinterpret_cast<const __half2*>(y2 + at));
            const float2 zl = __half22float2(*reinterpret_cast<const __half2*>(y2 + plane_stride + at));
            y0 += zh.x + zl.x;
            y1 += zh.y + zl.y;
        }
        v[2 * i] = (xh.x + xl.x) + y0;
        v[2 * i + 1] = (xh.y + xl.y) + y1;
        sum += v[2 * i] + v[2 * i + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / H;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * PER; ++i) {
        const float d = v[i] - mean;
        sq = fmaf(d, d, sq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / H + eps);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = 2 * (lane + 32 * i);
        const float a = (v[2 * i] - mean) * rstd * gamma[c] + beta[c];
        const float b = (v[2 * i + 1] - mean) * rstd * gamma[c + 1] + beta[c + 1];
        uint32_t hi2, lo2;
        tc::split2_f16(a, b, hi2, lo2);
        *reinterpret_cast<uint32_t*>(x + base + c) = hi2;
        *reinterpret_cast<uint32_t*>(x + plane_stride + base + c) = lo2;
        if (out_f32) {
            out_f32[base + c] = a;
            out_f32[base + c + 1] = b;
        }
    }
}

std::vector<float> k_major(const HostTensor& w) {   // (O, I, K) -> [O][tap*I + i]
    const int64_t O = w.shape[0], I = w.shape[1], K = w.shape[2];
    std::vector<float> out((size_t)(O * I * K));
    for (int64_t o = 0; o < O; ++o)
        for (int64_t i = 0; i < I; ++i)
            for (int64_t t = 0; t < K; ++t)
                out[(size_t)((o * K + t) * I + i)] = w.data[(size_t)((o * I + i) * K + t)];
    return out;
}

}  // namespace

struct W2v2Layer {
    float *qkv_w, *qkv_b, *out_w, *out_b, *ln1_w, *ln1_b, *ff1_w, *ff1_b, *ff2_w, *ff2_b, *ln2_w, *ln2_b;
};

struct W2v2TcLayer {
    TcWeight qkv, out, ff1, ff2;
};

struct W2v2Weights {
    void* blob = nullptr;
    // split-fp16 planes of the encoder projections for the tcgen05 GEMMs
    void* tc_blob = nullptr;
    float* tc_scales = nullptr;
    W2v2TcLayer tc[kLayers];
    TcWeight tc_conv[kNumConv];   // [1..6]: conv layers as tap GEMMs
    TcWeight tc_pos;              // positional conv: [16 groups x 64 rows (48 + zero pad)][128 taps x 48]
    float* conv_w[kNumConv];
    float *gn_w, *gn_b, *fp_ln_w, *fp_ln_b, *fp_w, *fp_b, *pos_w, *pos_b, *enc_ln_w, *enc_ln_b, *zero_bias;
    W2v2Layer layers[kLayers];
};

static std::vector<std::pair<std::string, std::vector<int64_t>>> w2v2_expected_keys() {
    std::vector<std::pair<std::string, std::vector<int64_t>>> keys;
    int c_in = 1;
    for (int i = 0; i < kNumConv; ++i) {
        keys.push_back({"feature_extractor.conv_layers." + std::to_string(i) + ".conv.weight",
                        {kConvDim, c_in, kConvKernel[i]}});
        c_in = kConvDim;
    }
    keys.push_back({"feature_extractor.conv_layers.0.layer_norm.weight", {kConvDim}});
    keys.push_back({"feature_extractor.conv_layers.0.layer_norm.bias", {kConvDim}});
    keys.push_back({"feature_projection.layer_norm.weight", {kConvDim}});
    keys.push_back({"feature_projection.layer_norm.bias", {kConvDim}});
    keys.push_back({"feature_projection.projection.weight", {kHidden, kConvDim}});
    keys.push_back({"feature_projection.projection.bias", {kHidden}});
    keys.push_back({"encoder.pos_conv_embed.conv.bias", {kHidden}});
    keys.push_back({"encoder.pos_conv_embed.conv.parametrizations.weight.original0", {1, 1, kPosKernel}});
    keys.push_back({"encoder.pos_conv_embed.conv.parametrizations.weight.original1",
                    {kHidden, kPosPer, kPosKernel}});
    keys.push_back({"encoder.layer_norm.weight", {kHidden}});
    keys.push_back({"encoder.layer_norm.bias", {kHidden}});
    for (int l = 0; l < kLayers; ++l) {
        const std::string p = "encoder.layers." + std::to_string(l) + ".";
        for (const char* name : {"q_proj", "k_proj", "v_proj", "out_proj"}) {
            keys.push_back({p + "attention." + name + ".weight", {kHidden, kHidden}});
            keys.push_back({p + "attention." + name + ".bias", {kHidden}});
        }
        keys.push_back({p + "layer_norm.weight", {kHidden}});
        keys.push_back({p + "layer_norm.bias", {kHidden}});
        keys.push_back({p + "feed_forward.intermediate_dense.weight", {kFfn, kHidden}});
        keys.push_back({p + "feed_forward.intermediate_dense.bias", {kFfn}});
        keys.push_back({p + "feed_forward.output_dense.weight", {kHidden, kFfn}});
        keys.push_back({p + "feed_forward.output_dense.bias", {kHidden}});
        keys.push_back({p + "final_layer_norm.weight", {kHidden}});
        keys.push_back({p + "final_layer_norm.bias", {kHidden}});
    }
    return keys;
}

int w2v2_accepts_key(const std::string& key, const std::vector<int64_t>& shape) {
    for (const auto& spec : w2v2_expected_keys()) {
        if (spec.first != key) continue;
        if (spec.second != shape) {
            set_error("size mismatch for w2v2.%s", key.c_str());
            return PPGS_E_INVALID;
        }
        return PPGS_OK;
    }
    if (key == "masked_spec_embed") return 1;   // present in HF checkpoints, unused in eval
    set_error("unexpected key in state_dict: w2v2.%s", key.c_str());
    return PPGS_E_INVALID;
}

void w2v2_free(ppgs_engine* e) {
    if (!e->w2v2) return;
    cudaFree(e->w2v2->blob);
    cudaFree(e->w2v2->tc_blob);
    cudaFree(e->w2v2->tc_scales);
    delete e->w2v2;
    e->w2v2 = nullptr;
}

int w2v2_finalize(ppgs_engine* e) {
    for (const auto& spec : w2v2_expected_keys())
        if (!e->w2v2_host.count(spec.first)) {
            set_error("missing key in state_dict: w2v2.%s", spec.first.c_str());
            return PPGS_E_STATE;
        }
    std::vector<float> host;
    std::vector<std::pair<float**, size_t>> slots;
    auto put = [&](float** dst, const std::vector<float>& data) {
        slots.push_back({dst, host.size()});
        host.insert(host.end(), data.begin(), data.end());
        host.resize((host.size() + 63) & ~size_t(63), 0.f);   // 256-byte aligned
    };
    auto raw = [&](float** dst, const std::string& key) { put(dst, e->w2v2_host.at(key).data); };
    w2v2_free(e);
    W2v2Weights* w = new W2v2Weights();
    e->w2v2 = w;
    for (int i = 0; i < kNumConv; ++i) {
        const HostTensor& t = e->w2v2_host.at("feature_extractor.conv_layers." + std::to_string(i) + ".conv.weight");
        if (i == 0) put(&w->conv_w[0], t.data);        // [512][1][10] is already [512][10]
        else put(&w->conv_w[i], k_major(t));
    }
    raw(&w->gn_w, "feature_extractor.conv_layers.0.layer_norm.weight");
    raw(&w->gn_b, "feature_extractor.conv_layers.0.layer_norm.bias");
    raw(&w->fp_ln_w, "feature_projection.layer_norm.weight");
    raw(&w->fp_ln_b, "feature_projection.layer_norm.bias");
    raw(&w->fp_w, "feature_projection.projection.weight");
    raw(&w->fp_b, "feature_projection.projection.bias");
    HostTensor pos_tc;   // folded positional-conv weight, one 64-row block per group
    {   // fold the weight norm (dim=2): w = v * g / ||v||, norm over (out, in) per tap; then
        // per group [48 out][tap*48 + in]
        const HostTensor& g = e->w2v2_host.at("encoder.pos_conv_embed.conv.parametrizations.weight.original0");
        const HostTensor& v = e->w2v2_host.at("encoder.pos_conv_embed.conv.parametrizations.weight.original1");
        std::vector<double> norm(kPosKernel, 0.0);
        for (int o = 0; o < kHidden; ++o)
            for (int i = 0; i < kPosPer; ++i)
                for (int t = 0; t < kPosKernel; ++t) {
                    const double x = v.data[(size_t)((o * kPosPer + i) * kPosKernel + t)];
                    norm[t] += x * x;
                }
        std::vector<float> packed((size_t)kHidden * kPosPer * kPosKernel);
        for (int o = 0; o < kHidden; ++o)
            for (int i = 0; i < kPosPer; ++i)
                for (int t = 0; t < kPosKernel; ++t) {
                    const float scale = g.data[t] / (float)sqrt(norm[t]);
                    packed[(size_t)o * kPosPer * kPosKernel + (size_t)t * kPosPer + i] =
                        v.data[(size_t)((o * kPosPer + i) * kPosKernel + t)] * scale;
                }
        put(&w->pos_w, packed);
        pos_tc.shape = {kPosGroups * 64, (int64_t)kPosPer * kPosKernel};
        pos_tc.data.assign((size_t)kPosGroups * 64 * kPosPer * kPosKernel, 0.f);
        for (int g = 0; g < kPosGroups; ++g)
            for (int o = 0; o < kPosPer; ++o)
                std::copy_n(packed.begin() + (size_t)(g * kPosPer + o) * kPosPer * kPosKernel,
                            (size_t)kPosPer * kPosKernel,
                            pos_tc.data.begin() + (size_t)(g * 64 + o) * kPosPer * kPosKernel);
    }
    raw(&w->pos_b, "encoder.pos_conv_embed.conv.bias");
    raw(&w->enc_ln_w, "encoder.layer_norm.weight");
    raw(&w->enc_ln_b, "encoder.layer_norm.bias");
    put(&w->zero_bias, std::vector<float>(kFfn, 0.f));
    for (int l = 0; l < kLayers; ++l) {
        const std::string p = "encoder.layers." + std::to_string(l) + ".";
        W2v2Layer& L = w->layers[l];
        std::vector<float> qkv_w, qkv_b;
        for (const char* name : {"q_proj", "k_proj", "v_proj"}) {
            const HostTensor& tw = e->w2v2_host.at(p + "attention." + name + ".weight");
            const HostTensor& tb = e->w2v2_host.at(p + "attention." + name + ".bias");
            qkv_w.insert(qkv_w.end(), tw.data.begin(), tw.data.end());
            qkv_b.insert(qkv_b.end(), tb.data.begin(), tb.data.end());
        }
        put(&L.qkv_w, qkv_w);
        put(&L.qkv_b, qkv_b);
        raw(&L.out_w, p + "attention.out_proj.weight");
        raw(&L.out_b, p + "attention.out_proj.bias");
        raw(&L.ln1_w, p + "layer_norm.weight");
        raw(&L.ln1_b, p + "layer_norm.bias");
        raw(&L.ff1_w, p + "feed_forward.intermediate_dense.weight");
        raw(&L.ff1_b, p + "feed_forward.intermediate_dense.bias");
        raw(&L.ff2_w, p + "feed_forward.output_dense.weight");
        raw(&L.ff2_b, p + "feed_forward.output_dense.bias");
        raw(&L.ln2_w, p + "final_layer_norm.weight");
        raw(&L.ln2_b, p + "final_layer_norm.bias");
    }
    PPGS_CUDA(cudaMalloc(&w->blob, host.size() * sizeof(float)));
    PPGS_CUDA(cudaMemcpy(w->blob, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    for (auto& slot : slots) *slot.first = static_cast<float*>(w->blob) + slot.second;

    // tensor-core copies of the encoder projections: [2][N][K] planes + 1/scale each
    {
        std::vector<__half> planes_all;
        std::vector<float> scales;
        struct Slot { TcWeight* w; size_t offset; int N, K, taps; };
        std::vector<Slot> tc_slots;
        auto pack = [&](TcWeight* dst, const HostTensor& t) {
            std::vector<__half> planes;
            scales.push_back(pack_planes(&t, planes));
            tc_slots.push_back({dst, planes_all.size(), (int)t.shape[0], (int)t.shape[1],
                                t.shape.size() == 3 ? (int)t.shape[2] : 1});
            planes_all.insert(planes_all.end(), planes.begin(), planes.end());
            planes_all.resize((planes_all.size() + 127) & ~size_t(127), __float2half_rn(0.f));
        };
        for (int i = 1; i < kNumConv; ++i)
            pack(&w->tc_conv[i],
                 e->w2v2_host.at("feature_extractor.conv_layers." + std::to_string(i) + ".conv.weight"));
        pack(&w->tc_pos, pos_tc);
        for (int l = 0; l < kLayers; ++l) {
            const std::string p = "encoder.layers." + std::to_string(l) + ".";
            HostTensor qkv;
            qkv.shape = {3 * kHidden, kHidden};
            for (const char* name : {"q_proj", "k_proj", "v_proj"}) {
                const HostTensor& tw = e->w2v2_host.at(p + "attention." + name + ".weight");
                qkv.data.insert(qkv.data.end(), tw.data.begin(), tw.data.end());
            }
            pack(&w->tc[l].qkv, qkv);
            pack(&w->tc[l].out, e->w2v2_host.at(p + "attention.out_proj.weight"));
            pack(&w->tc[l].ff1, e->w2v2_host.at(p + "feed_forward.intermediate_dense.weight"));
            pack(&w->tc[l].ff2, e->w2v2_host.at(p + "feed_forward.output_dense.weight"));
        }
        PPGS_CUDA(cudaMalloc(&w->tc_blob, planes_all.size() * sizeof(__half)));
        PPGS_CUDA(cudaMemcpy(w->tc_blob, planes_all.data(), planes_all.size() * sizeof(__half),
                             cudaMemcpyHostToDevice));
        PPGS_CUDA(cudaMalloc(&w->tc_scales, scales.size() * sizeof(float)));
        PPGS_CUDA(cudaMemcpy(w->tc_scales, scales.data(), scales.size() * sizeof(float),
                             cudaMemcpyHostToDevice));
        for (size_t i = 0; i < tc_slots.size(); ++i) {
            TcWeight& tw = *tc_slots[i].w;
            tw.planes = static_cast<__half*>(w->tc_blob) + tc_slots[i].offset;
            tw.inv_scale = w->tc_scales + i;
            tw.N = tc_slots[i].N;
            tw.C = tc_slots[i].K;
            tw.taps = tc_slots[i].taps;
            PPGS_CHECK(build_weight_map(tw));
        }
        PPGS_CHECK(ensure_status_word(e));
    }
    e->w2v2_host.clear();
    return PPGS_OK;
}

int w2v2fb_forward(ppgs_engine* e, const float* audio, int batch, int64_t samples, int64_t stride,
                   const int64_t* lengths, __half* out, cudaStream_t stream) {
    if (!e->w2v2) {
        set_error("w2v2fb: no wav2vec2 weights loaded (ppgs_engine_set_weight(\"w2v2.*\") + "
                  "ppgs_w2v2_finalize)");
        return PPGS_E_STATE;
    }
    const W2v2Weights& w = *e->w2v2;
    const int frames = (int)(samples / kHopSamples);
    int T[kNumConv];
    int64_t t = samples + 2 * kW2v2Pad;
    for (int l = 0; l < kNumConv; ++l) {
        t = (t - kConvKernel[l]) / kConvStride[l] + 1;
        if (t <= 0) {
            set_error("w2v2fb: audio too short (%lld samples)", (long long)samples);
            return PPGS_E_INVALID;
        }
        T[l] = (int)t;
    }
    const int T6 = T[6];
    // Pitches.  Conv layers: Q_{l-1} = 2 Q_l with Q_l >= T_l, so a stride-2 convolution is a
    // GEMM over overlapping A rows.  Encoder: P >= T6, multiple of 128; the positional
    // convolution works at Pg = P + 128 (zero halo rows between utterances).  M = encoder
    // rows, padded to a multiple of 256 (CTA-pair tiles) with dummy rows.
    int64_t Q[kNumConv];
    Q[6] = 1;
    for (int l = 0; l < kNumConv; ++l) Q[6] = std::max<int64_t>(Q[6], ((int64_t)T[l] + (1 << (6 - l)) - 1) >> (6 - l));
    for (int l = 5; l >= 0; --l) Q[l] = 2 * Q[l + 1];
    const int64_t P = ((int64_t)T6 + 127) / 128 * 128, Pg = P + 128;
    const int64_t M = ((int64_t)batch * P + 255) / 256 * 256, Mg = (int64_t)batch * Pg;
    if ((int64_t)batch * Q[0] + 1024 > INT32_MAX || Mg > INT32_MAX) {
        set_error("w2v2fb: batch too large");
        return PPGS_E_TOO_LARGE;
    }
    std::vector<int> out_len(batch);
    std::vector<SeqInfo> seqs(batch);
    for (int b = 0; b < batch; ++b) {
        int64_t len = (lengths ? lengths[b] : samples) + 2 * kW2v2Pad;
        for (int l = 0; l < kNumConv; ++l) len = (len - kConvKernel[l]) / kConvStride[l] + 1;
        out_len[b] = (int)std::max<int64_t>(std::min<int64_t>(len, T6), 0);
        SeqInfo s{};
        s.row0 = (int)(b * P);
        s.tensor_len = T6;
        s.valid_len = out_len[b];
        s.batch = b;
        seqs[b] = s;
    }
    static const bool use_tc = [] { const char* v = getenv("PPGS_B200_W2V2_TC"); return !v || atoi(v) != 0; }();
    static const int split_acc = [] { const char* v = getenv("PPGS_B200_W2V2_SPLIT_ACC"); return v ? atoi(v) : 1; }();

    // workspace: two conv activation buffers (fp32 rows, or split-fp16 planes: same bytes
    // per element) + encoder buffers (fp32).  512 slack rows: GEMM tiles round M up.
    Carver c;
    // (the encoder's split planes are carved out of the same two buffers once the convs are done)
    const size_t plane_bytes = (size_t)M * (3 * kHidden + 3 * kHidden + kFfn) * 4 + 5 * 1024;
    const size_t act_bytes = std::max({((size_t)batch * Q[0] + 512) * kConvDim * 4,
                                       (size_t)M * kConvDim * 4, (plane_bytes + 1) / 2});
    const size_t o_act0 = c.take(act_bytes);
    const size_t o_act1 = c.take(act_bytes);
    const size_t o_sums = c.take((size_t)batch * kConvDim * 2 * sizeof(double));
    const size_t o_h = c.take((size_t)M * kHidden * 4);
    const size_t o_y = c.take((size_t)std::max(M, Mg) * kHidden * 4);
    const size_t o_qkv = c.take((size_t)M * 3 * kHidden * 4);
    const size_t o_ff = c.take((size_t)M * kFfn * 4);
    const size_t xg_bytes = (size_t)kPosGroups * (Mg + 128) * kPosPer * 4 + 16 * kConvDim * 4;
    const size_t o_xg = c.take(xg_bytes);
    const size_t o_seqs = c.take(batch * sizeof(SeqInfo));
    const size_t o_len = c.take(batch * sizeof(int));
    PPGS_CHECK(ensure_workspace(e, c.off));
    e->cached_plan_dev = nullptr;   // the PPG transformer's plan tables are overwritten
    char* ws = static_cast<char*>(e->workspace);
    float* act[2] = {reinterpret_cast<float*>(ws + o_act0), reinterpret_cast<float*>(ws + o_act1)};
    double* sums = reinterpret_cast<double*>(ws + o_sums);
    float* h = reinterpret_cast<float*>(ws + o_h);
    float* y = reinterpret_cast<float*>(ws + o_y);
    float* qkv = reinterpret_cast<float*>(ws + o_qkv);
    float* ff = reinterpret_cast<float*>(ws + o_ff);
    float* xg = reinterpret_cast<float*>(ws + o_xg);
    SeqInfo* seqs_dev = reinterpret_cast<SeqInfo*>(ws + o_seqs);
    int* len_dev = reinterpret_cast<int*>(ws + o_len);
    // small tables: synchronous pageable copies are fine here (two tiny transfers per call)
    PPGS_CUDA(cudaMemcpyAsync(seqs_dev, seqs.data(), batch * sizeof(SeqInfo), cudaMemcpyHostToDevice, stream));
    PPGS_CUDA(cudaMemcpyAsync(len_dev, out_len.data(), batch * sizeof(int), cudaMemcpyHostToDevice, stream));
    PPGS_CUDA(cudaStreamSynchronize(stream));   // host vectors go out of scope

    // ---- feature encoder: conv0 + GroupNorm + GELU on the CUDA cores, then six stride-2
    // convolutions as GEMMs (tensor cores: taps = k-blocks, A boxes take every 2nd row)
    // fp32 path: junk rows of the conv activations must be finite.  The tensor-core path writes
    // every row a TMA load can see (maps are bounded by the row count, stores are clipped)
    if (!use_tc) PPGS_CUDA(cudaMemsetAsync(act[0], 0, act_bytes, stream));
    PPGS_CUDA(cudaMemsetAsync(sums, 0, (size_t)batch * kConvDim * 2 * sizeof(double), stream));
    if (use_tc) {
        LaunchScope scope(e, "w2v2_conv0_stats", stream);
        conv0_stats_kernel<<<dim3((T[0] + 255) / 256, batch), 512, 0, stream>>>(audio, stride, (int)samples, T[0],
                                                                               w.conv_w[0], sums);
    } else {
        {
            LaunchScope scope(e, "w2v2_conv0", stream);
            conv0_kernel<<<dim3((T[0] + 31) / 32, batch), 256, 0, stream>>>(audio, stride, (int)samples, T[0],
                                                                              Q[0], w.conv_w[0], act[0]);
        }
        LaunchScope scope(e, "w2v2_groupnorm_stats", stream);
        groupnorm_stats_kernel<<<dim3((T[0] + 255) / 256, batch), 512, 0, stream>>>(act[0], T[0], Q[0], sums);
    }
    float* normed;   // LayerNorm(512) of the conv features, encoder pitch
    if (use_tc) {
        using namespace tc;
        // planes of layer l: [2][B*Q_l (+ slack)][512] fp16, alternating between the buffers
        __half* bufs[2] = {reinterpret_cast<__half*>(act[0]), reinterpret_cast<__half*>(act[1])};
        const int64_t rows0 = (int64_t)batch * Q[0];
        {
            LaunchScope scope(e, "w2v2_conv0_groupnorm_gelu", stream);
            conv0_groupnorm_gelu_planes_kernel<<<dim3((unsigned)((Q[0] + 63) / 64), batch), 256, 0, stream>>>(
                audio, stride, (int)samples, T[0], Q[0], rows0 * kConvDim, w.conv_w[0], sums, w.gn_w, w.gn_b,
                1e-5f, bufs[1]);
        }
        PPGS_CUDA(cudaGetLastError());
        int cur = 1;
        for (int l = 1; l < kNumConv; ++l) {
            const int64_t rows_in = (int64_t)batch * Q[l - 1], rows_out = (int64_t)batch * Q[l];
            TcWeight& wt = e->w2v2->tc_conv[l];
            CUtensorMap map_a, map_out;
            PPGS_CHECK(make_plane_map(&map_a, bufs[cur], false, kConvDim, rows_in, 1, 2, kConvDim, 0,
                                      (uint64_t)rows_in * kConvDim, 128, 2, 2));
            PPGS_CHECK(make_store_map(&map_out, bufs[cur ^ 1], kConvDim, rows_out,
                                      (uint64_t)rows_out * kConvDim));
            GemmParams p;
            p.m_tiles = (int)((rows_out + 255) / 256) * 2;
            p.n_tiles = kConvDim / 256;
            p.cblocks = kConvDim / 64;
            p.taps = kConvKernel[l];
            p.half = 0;
            p.row_mul = 2;
            p.a_planes = 2;
            p.b_planes = 2;
            p.pair = 1;
            p.N = kConvDim;
            p.scale = wt.inv_scale;
            p.bias = w.zero_bias;
            p.relu = 2;   // GELU
            p.status = e->status_dev;
            p.split_acc = split_acc;
            PPGS_CHECK(launch_gemm_tc(e, "w2v2_tc_conv", 256, kEpiPlanes, map_a, wt.maps[1].bn128, &map_out,
                                      p, stream));
            cur ^= 1;
        }
        normed = reinterpret_cast<float*>(bufs[cur ^ 1]);
        LaunchScope scope(e, "w2v2_layernorm", stream);
        feature_layernorm_kernel<true><<<(unsigned)((M + 7) / 8), 256, 0, stream>>>(
            bufs[cur], (int64_t)batch * Q[6] * kConvDim, Q[6], P, T6, batch, w.fp_ln_w, w.fp_ln_b, 1e-5f,
            (int)M, normed);
    } else {
        {
            LaunchScope scope(e, "w2v2_groupnorm_gelu", stream);
            groupnorm_gelu_kernel<<<dim3((T[0] + 63) / 64, batch), 512, 0, stream>>>(
                act[0], T[0], Q[0], sums, w.gn_w, w.gn_b, 1e-5f);
        }
        PPGS_CUDA(cudaGetLastError());
        int cur = 0;
        for (int l = 1; l < kNumConv; ++l) {
            SgemmArgs a{};
            a.A = act[cur];
            a.lda = (int64_t)kConvStride[l] * kConvDim;
            a.B = w.conv_w[l];
            a.bias = w.zero_bias;
            a.out = act[cur ^ 1];
            a.ldo = kConvDim;
            a.M = (int)(((int64_t)batch * Q[l] + 127) / 128 * 128);   // tail rows land in the slack
            a.N = kConvDim;
            a.K = kConvKernel[l] * kConvDim;
            PPGS_CHECK(launch_sgemm_any(e, "w2v2_conv_gemm", EPI_BIAS_GELU, a, stream));
            cur ^= 1;
        }
        normed = act[cur ^ 1];
        LaunchScope scope(e, "w2v2_layernorm", stream);
        feature_layernorm_kernel<false><<<(unsigned)((M + 7) / 8), 256, 0, stream>>>(
            act[cur], 0, Q[6], P, T6, batch, w.fp_ln_w, w.fp_ln_b, 1e-5f, (int)M, normed);
    }
    PPGS_CUDA(cudaGetLastError());

    // ---- feature projection + padding mask
    {
        SgemmArgs a{};
        a.A = normed; a.lda = kConvDim; a.B = w.fp_w; a.bias = w.fp_b; a.out = h; a.ldo = kHidden;
        a.M = (int)M; a.N = kHidden; a.K = kConvDim;
        PPGS_CHECK(launch_sgemm_any(e, "w2v2_projection", EPI_BIAS, a, stream));
    }
    {
        LaunchScope scope(e, "w2v2_mask", stream);
        mask_rows_kernel<<<(unsigned)M, 256, 0, stream>>>(h, P, len_dev, kHidden, batch);
    }

    // ---- positional convolution (grouped, k=128) + LayerNorm.  Group g of output row m is a
    // dot product over 128 x 48 contiguous values of the guarded group-major copy starting at
    // row m: a GEMM whose A rows overlap (row stride 48 elements, K = 6144).
    PPGS_CUDA(cudaMemsetAsync(xg, 0, xg_bytes, stream));
    if (use_tc) {
        using namespace tc;
        __half* xgh = reinterpret_cast<__half*>(xg);
        {
            LaunchScope scope(e, "w2v2_group_major", stream);
            group_major_planes_kernel<<<(unsigned)((int64_t)batch * P), 256, 0, stream>>>(h, P, Pg, Mg, xgh);
        }
        const uint64_t group_rows = (uint64_t)(Mg + 128);
        CUtensorMap map_a;
        PPGS_CHECK(make_plane_map(&map_a, xgh, false, (uint64_t)kPosKernel * kPosPer,
                                  kPosGroups * group_rows - 128, 1, 2, kPosPer, 0,
                                  kPosGroups * group_rows * kPosPer, 128, 2));
        TcWeight& wt = e->w2v2->tc_pos;
        GemmParams p;
        p.m_tiles = (int)(Mg / 128);
        p.n_tiles = kPosGroups;
        p.cblocks = kPosKernel * kPosPer / 64;
        p.a_group_rows = (int)group_rows;
        p.group_cols = kPosPer;
        p.a_planes = 2;
        p.b_planes = 2;
        p.N = kHidden;
        p.scale = wt.inv_scale;
        p.bias = w.pos_b;
        p.relu = 2;   // GELU
        p.out_f32 = y;
        p.ld_f32 = kHidden;
        p.status = e->status_dev;
        PPGS_CHECK(launch_gemm_tc(e, "w2v2_tc_pos_conv", 64, kEpiF32, map_a, wt.maps[1].bn64, nullptr, p, stream));
    } else {
        {
            LaunchScope scope(e, "w2v2_group_major", stream);
            group_major_kernel<<<(unsigned)((int64_t)batch * P), 256, 0, stream>>>(h, P, Pg, Mg, xg);
        }
        for (int g = 0; g < kPosGroups; ++g) {
            SgemmArgs a{};
            a.A = xg + (int64_t)g * (Mg + 128) * kPosPer;     // output row m <- rows m .. m+127 of the guarded copy
            a.lda = kPosPer;
            a.B = w.pos_w + (int64_t)g * kPosPer * kPosPer * kPosKernel;
            a.bias = w.pos_b + g * kPosPer;
            a.out = y + g * kPosPer;
            a.ldo = kHidden;
            a.M = (int)Mg; a.N = kPosPer; a.K = kPosKernel * kPosPer;
            PPGS_CHECK(launch_sgemm_any(e, "w2v2_pos_conv_gemm", EPI_BIAS_GELU, a, stream));
        }
    }
    {
        LaunchScope scope(e, "w2v2_layernorm", stream);
        add_layernorm_kernel<kHidden><<<(unsigned)((M + 7) / 8), 256, 0, stream>>>(
            h, y, w.enc_ln_w, w.enc_ln_b, 1e-5f, (int)M, h, P, Pg, batch);
    }
    PPGS_CUDA(cudaGetLastError());

    // ---- 12 post-LN encoder layers.  Default: projections / FFN on the tensor cores
    // (split-fp16 3-pass CTA-pair GEMMs), attention and LayerNorm over the split planes on
    // the CUDA cores; PPGS_B200_W2V2_TC=0 keeps the all-fp32 CUDA-core layers below.
    if (use_tc) {
        using namespace tc;
        Carver pc;
        const size_t o_xh = pc.take((size_t)2 * M * kHidden * 2);
        const size_t o_yh = pc.take((size_t)2 * M * kHidden * 2);
        const size_t o_ah = pc.take((size_t)2 * M * kHidden * 2);
        const size_t o_qh = pc.take((size_t)2 * M * 3 * kHidden * 2);
        const size_t o_fh = pc.take((size_t)2 * M * kFfn * 2);
        // the conv activation buffers are dead by now: carve the planes out of them
        if (pc.off > o_h) {
            set_error("w2v2fb: internal workspace layout error");
            return PPGS_E_STATE;
        }
        __half* xh = reinterpret_cast<__half*>(ws + o_xh);
        __half* yh = reinterpret_cast<__half*>(ws + o_yh);
        __half* ah = reinterpret_cast<__half*>(ws + o_ah);
        __half* qh = reinterpret_cast<__half*>(ws + o_qh);
        __half* fh = reinterpret_cast<__half*>(ws + o_fh);
        CUtensorMap a_x, a_att, a_ff, s_qkv, s_y, s_ff;
        PPGS_CHECK(make_plane_map(&a_x, xh, false, kHidden, M, 1, 2, kHidden, 0, (uint64_t)M * kHidden, 128, 2));
        PPGS_CHECK(make_plane_map(&a_att, ah, false, kHidden, M, 1, 2, kHidden, 0, (uint64_t)M * kHidden, 128, 2));
        PPGS_CHECK(make_plane_map(&a_ff, fh, false, kFfn, M, 1, 2, kFfn, 0, (uint64_t)M * kFfn, 128, 2));
        PPGS_CHECK(make_store_map(&s_qkv, qh, 3 * kHidden, M, (uint64_t)M * 3 * kHidden));
        PPGS_CHECK(make_store_map(&s_y, yh, kHidden, M, (uint64_t)M * kHidden));
        PPGS_CHECK(make_store_map(&s_ff, fh, kFfn, M, (uint64_t)M * kFfn));
        PPGS_CUDA(cudaMemsetAsync(ah, 0, (size_t)2 * M * kHidden * 2, stream));   // rows no sequence owns
        {
            const int64_t count = M * kHidden;
            LaunchScope scope(e, "w2v2_to_planes", stream);
            to_planes_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(h, count, xh);
        }
        // K split (`parts` launches over K ranges, partial sums added in fp32 by the LayerNorm
        // pass): tcgen05 accumulation TRUNCATES, so an accumulator's error grows with its number
        // of MMA steps; the K = 3072 projection is the longest chain of the encoder
        auto gemm = [&](const char* name, const CUtensorMap& a, TcWeight& wt, const CUtensorMap& out_map,
                        const float* bias, int act, int reverse, int part = 0, int parts = 1) -> int {
            GemmParams p;
            p.reverse = e->serpentine ? reverse : 0;   // serpentine row order across the kernels
            p.m_tiles = (int)(M / 128);
            p.n_tiles = wt.N / 256;
            p.cblocks = wt.C / 64 / parts;
            p.cb0 = part * p.cblocks;
            p.a_planes = 2;
            p.b_planes = 2;
            p.pair = 1;
            p.N = wt.N;
            p.scale = wt.inv_scale;
            p.bias = bias;
            p.relu = act;
            p.status = e->status_dev;
            p.split_acc = split_acc;
            return launch_gemm_tc(e, name, 256, kEpiPlanes, a, wt.maps[1].bn128, &out_map, p, stream);
        };
        auto add_ln = [&](const float* g, const float* b2, float* f32, const __half* y2 = nullptr) {
            LaunchScope scope(e, "w2v2_add_layernorm_planes", stream);
            add_layernorm_planes_kernel<kHidden><<<(unsigned)((M + 7) / 8), 256, 0, stream>>>(
                xh, yh, y2, M * kHidden, g, b2, 1e-5f, (int)M, f32);
        };
        // second partial sum of the K-split FFN projection: the QKV planes are dead by then
        CUtensorMap s_y2;
        PPGS_CHECK(make_store_map(&s_y2, qh, kHidden, M, (uint64_t)M * kHidden));
        const int ffn2_parts = split_acc ? 2 : 1;
        for (int l = 0; l < kLayers; ++l) {
            const W2v2Layer& L = w.layers[l];
            W2v2TcLayer& T = e->w2v2->tc[l];
            // row direction per kernel: each starts where its predecessor (GEMM, attention or the
            // ascending LayerNorm pass) ended
            PPGS_CHECK(gemm("w2v2_tc_qkv", a_x, T.qkv, s_qkv, L.qkv_b, 0, 1));
            PPGS_CHECK(launch_attention_any(e, qh, ah, (int)M, kHidden, kHeads, (int)P, batch, seqs_dev, 0,
                                            2, stream));
            PPGS_CHECK(gemm("w2v2_tc_out_proj", a_att, T.out, s_y, L.out_b, 0, 1));
            add_ln(L.ln1_w, L.ln1_b, nullptr);
            PPGS_CHECK(gemm("w2v2_tc_ffn1", a_x, T.ff1, s_ff, L.ff1_b, 2, 1));
            PPGS_CHECK(gemm("w2v2_tc_ffn2", a_ff, T.ff2, s_y, L.ff2_b, 0, 0, 0, ffn2_parts));
            if (ffn2_parts == 2) PPGS_CHECK(gemm("w2v2_tc_ffn2", a_ff, T.ff2, s_y2, w.zero_bias, 0, 1, 1, 2));
            add_ln(L.ln2_w, L.ln2_b, l == kLayers - 1 ? h : nullptr, ffn2_parts == 2 ? qh : nullptr);
            PPGS_CUDA(cudaGetLastError());
        }
    } else
    for (int l = 0; l < kLayers; ++l) {
        const W2v2Layer& L = w.layers[l];
        SgemmArgs a{};
        a.M = (int)M;
        a.A = h; a.lda = kHidden; a.B = L.qkv_w; a.bias = L.qkv_b; a.out = qkv; a.ldo = 3 * kHidden;
        a.N = 3 * kHidden; a.K = kHidden;
        PPGS_CHECK(launch_sgemm_any(e, "w2v2_qkv", EPI_BIAS, a, stream));
        PPGS_CHECK(launch_attention_fp32_any(e, kHidden / kHeads, qkv, kHidden, kHeads, (int)P, batch,
                                             seqs_dev, 0, y, stream));
        a.A = y; a.B = L.out_w; a.bias = L.out_b; a.out = qkv; a.ldo = kHidden; a.N = kHidden; a.res = h;
        PPGS_CHECK(launch_sgemm_any(e, "w2v2_out_proj", EPI_BIAS_RES, a, stream));
        {
            LaunchScope scope(e, "w2v2_layernorm", stream);
            add_layernorm_kernel<kHidden><<<(unsigned)((M + 7) / 8), 256, 0, stream>>>(
                qkv, nullptr, L.ln1_w, L.ln1_b, 1e-5f, (int)M, h);
        }
        a.A = h; a.B = L.ff1_w; a.bias = L.ff1_b; a.out = ff; a.ldo = kFfn; a.N = kFfn; a.K = kHidden;
        a.res = nullptr;
        PPGS_CHECK(launch_sgemm_any(e, "w2v2_ffn1", EPI_BIAS_GELU, a, stream));
        a.A = ff; a.lda = kFfn; a.B = L.ff2_w; a.bias = L.ff2_b; a.out = qkv; a.ldo = kHidden;
        a.N = kHidden; a.K = kFfn; a.res = h;
        PPGS_CHECK(launch_sgemm_any(e, "w2v2_ffn2", EPI_BIAS_RES, a, stream));
        {
            LaunchScope scope(e, "w2v2_layernorm", stream);
            add_layernorm_kernel<kHidden><<<(unsigned)((M + 7) / 8), 256, 0, stream>>>(
                qkv, nullptr, L.ln2_w, L.ln2_b, 1e-5f, (int)M, h);
        }
        PPGS_CUDA(cudaGetLastError());
    }

    // ---- nearest upsample to the PPG frame rate, fp16 (w2v2fb/core.py:70-75)
    {
        LaunchScope scope(e, "w2v2_upsample", stream);
        const float scale = (float)T6 / (float)frames;
        upsample_kernel<<<dim3((frames + 31) / 32, kHidden / 32, batch), dim3(32, 8), 0, stream>>>(
            h, P, T6, frames, scale, out);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

}  // namespace ppgs
