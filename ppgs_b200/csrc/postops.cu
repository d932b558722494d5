// Posteriorgram post-processing on the device — the step right after the hot path
// (SURVEY.md §8 f3), so that distances / sparse or time-stretched PPGs do not need a D2H round
// trip:
//   ppgs.distance            ppgs/core.py:399-469   similarity-weighted Jensen-Shannon distance
//   ppgs.interpolate         ppgs/core.py:475-496   linear interpolation of two PPGs
//   ppgs.sparsify            ppgs/core.py:504-543   constant / percentile / top-k + renormalise
//   ppgs.edit.grid.sample    ppgs/edit/grid.py:13-50 float-index gather with linear interpolation
// One warp per frame: a lane holds phonemes `lane` and `lane + 32` (len(ppgs.PHONEMES) = 40).
// Compiled with --fmad=false: the reference evaluates products and sums separately.
#include <float.h>

#include "common.cuh"

namespace ppgs {

constexpr int kMaxPhonemes = 64;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// weights[i][j] = similarity[j][i] ** exponent   (`similarity_matrix.T ** exponent`, core.py:443)
__global__ void similarity_weights_kernel(const float* __restrict__ similarity, int P, float exponent,
                                          float* __restrict__ weights) {
    const int i = blockIdx.x, j = threadIdx.x;
    if (i < P && j < P) weights[i * P + j] = powf(similarity[j * P + i], exponent);
}

// torch.xlogy(x, x) - x * m  (F.kl_div(input=m, target=x, reduction='none'))
__device__ __forceinline__ float kl_term(float x, float m) {
    const float xlogx = x == 0.f ? 0.f : x * logf(x);
    return xlogx - x * m;
}

__global__ void __launch_bounds__(256)
distance_kernel(const float* __restrict__ X, const float* __restrict__ Y, int P, int64_t T, int64_t ldx,
                int64_t ldy, const float* __restrict__ weights, float* __restrict__ out) {
    extern __shared__ float smem[];
    float* w = smem;                                   // [P][P] when normalising
    float* frame = smem + (weights ? P * P : 0);       // per warp: x[P], y[P]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (weights)
        for (int i = threadIdx.x; i < P * P; i += blockDim.x) w[i] = weights[i];
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (t >= T) return;
    float* fx = frame + warp * 2 * P;
    float* fy = fx + P;
    float x[2], y[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int p = lane + 32 * h;
        x[h] = y[h] = 0.f;
        if (p < P) {
            // torch.clamp(ppg, 1e-8, 1 - 1e-8)   (core.py:433-434; 1 - 1e-8 rounds to 1 in fp32)
            x[h] = fminf(fmaxf(X[p * ldx + t], 1e-8f), 1.f - 1e-8f);
            y[h] = fminf(fmaxf(Y[p * ldy + t], 1e-8f), 1.f - 1e-8f);
        }
    }
    if (weights) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int p = lane + 32 * h;
            if (p < P) {
                fx[p] = x[h];
                fy[p] = y[h];
            }
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int p = lane + 32 * h;
            if (p >= P) continue;
            float ax = 0.f, ay = 0.f;
            for (int j = 0; j < P; ++j) {
                ax += w[p * P + j] * fx[j];
                ay += w[p * P + j] * fy[j];
            }
            x[h] = ax;
            y[h] = ay;
        }
    }
    float total = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (lane + 32 * h >= P) continue;
        const float m = logf((x[h] + y[h]) / 2.f);
        const float kl = (kl_term(x[h], m) + kl_term(y[h], m)) / 2.f;
        total += sqrtf(fmaxf(kl, 0.f));   // average_kl[average_kl < 0] = 0; sqrt (core.py:458-460)
    }
    total = warp_sum(total);
    if (lane == 0) out[t] = total;
}

// deterministic sum of n floats by one block (pairwise inside the block)
__global__ void __launch_bounds__(1024)
reduce_sum_kernel(const float* __restrict__ in, int64_t n, float scale, float* __restrict__ out) {
    __shared__ float part[32];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += in[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = warp_sum(part[threadIdx.x]);
        if (threadIdx.x == 0) out[0] = acc * scale;
    }
}

// (1 - w) * x + w * y, w per frame or scalar
__global__ void interpolate_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                   const float* __restrict__ interp, float scalar, int64_t T, int64_t count,
                                   float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float w = interp ? interp[i % T] : scalar;
    out[i] = (1.f - w) * X[i] + w * Y[i];
}

// ppg (P, T), grid (G,) -> out (P, G)
__global__ void grid_sample_kernel(const float* __restrict__ ppg, int P, int64_t T, const float* __restrict__ grid,
                                   int64_t G, float* __restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const float v = grid[g];
    const float w = v - floorf(v);
    // i = searchsorted(arange(T), v, side='right') = #{k : k <= v}
    int64_t i = v < 0.f ? 0 : (int64_t)floorf(v) + 1;
    if (i > T) i = T;
    // index into the PPG padded by one replicated frame; -1 wraps to that last entry
    int64_t lo = i - 1, hi = i;
    if (lo < 0) lo = T;
    if (lo >= T) lo = T - 1;
    if (hi >= T) hi = T - 1;
    for (int p = 0; p < P; ++p)
        out[p * G + g] = (1.f - w) * ppg[p * T + lo] + w * ppg[p * T + hi];
}

enum { kSparsifyConstant = 0, kSparsifyPercentile = 1, kSparsifyTopk = 2 };

__global__ void __launch_bounds__(256)
sparsify_kernel(const float* __restrict__ ppg, int P, int64_t T, int64_t frames_total, int method,
                float threshold, float* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;   // (batch, frame) flattened
    if (f >= frames_total) return;
    const int64_t b = f / T, t = f - b * T;
    const float* src = ppg + b * P * T + t;
    float v[2];
    int rank[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int p = lane + 32 * h;
        v[h] = p < P ? src[(int64_t)p * T] : -FLT_MAX;
    }
    if (method != kSparsifyConstant) {
        // rank = number of phonemes that sort before this one (ties broken by index)
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
            for (int l = 0; l < 32; ++l) {
                const float other = __shfl_sync(0xffffffffu, v[h2], l);
                const int q = l + 32 * h2;
                if (q >= P) continue;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int p = lane + 32 * h;
                    rank[h] += (other < v[h] || (other == v[h] && q < p)) ? 1 : 0;
                }
            }
    }
    float thr = threshold;
    bool keep[2];
    if (method == kSparsifyPercentile) {
        // torch.quantile(..., interpolation='linear'): position q * (P - 1) in fp32, lerp of
        // the two neighbouring order statistics
        const float pos = threshold * (float)(P - 1);
        const int lo = (int)floorf(pos), hi = (int)ceilf(pos);
        const float frac = pos - (float)lo;
        float vlo = 0.f, vhi = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const bool valid = lane + 32 * h < P;
            const unsigned mlo = __ballot_sync(0xffffffffu, valid && rank[h] == lo);
            const unsigned mhi = __ballot_sync(0xffffffffu, valid && rank[h] == hi);
            const float clo = __shfl_sync(0xffffffffu, v[h], mlo ? __ffs(mlo) - 1 : 0);
            const float chi = __shfl_sync(0xffffffffu, v[h], mhi ? __ffs(mhi) - 1 : 0);
            if (mlo) vlo = clo;
            if (mhi) vhi = chi;
        }
        // at::lerp: a + w (b - a) for w < 0.5, b - (b - a)(1 - w) otherwise
        const float diff = vhi - vlo;
        thr = frac < 0.5f ? vlo + frac * diff : vhi - diff * (1.f - frac);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (method == kSparsifyTopk) keep[h] = rank[h] >= P - (int)threshold;
        else keep[h] = v[h] > thr;
    }
    // softmax(log(ppg + 1e-8)) over the phonemes (core.py:540)
    float l[2], mx = -FLT_MAX;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const bool valid = lane + 32 * h < P;
        l[h] = valid ? logf((keep[h] ? v[h] : 0.f) + 1e-8f) : -FLT_MAX;
        mx = fmaxf(mx, l[h]);
    }
    mx = warp_max(mx);
    float e[2], sum = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        e[h] = lane + 32 * h < P ? expf(l[h] - mx) : 0.f;
        sum += e[h];
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int p = lane + 32 * h;
        if (p < P) out[b * P * T + (int64_t)p * T + t] = e[h] / sum;
    }
}

struct PostGuard {
    int prev = -1;
    bool ok = true;
    explicit PostGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) {
            set_error("cannot select CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
            ok = false;
        }
    }
    ~PostGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace ppgs

using namespace ppgs;

#define PPGS_POST_ENTER(e)                    \
    if (!(e)) {                               \
        set_error("engine is NULL");          \
        return PPGS_E_INVALID;                \
    }                                         \
    PostGuard guard__((e)->device);           \
    if (!guard__.ok) return PPGS_E_CUDA

extern "C" {

int ppgs_ppg_distance(ppgs_engine* e, const float* x_dev, const float* y_dev, int phonemes, int64_t frames,
                      int64_t x_stride, int64_t y_stride, const float* similarity_dev, float exponent,
                      int reduction, float* out_dev, void* stream_) {
    PPGS_POST_ENTER(e);
    if (!x_dev || !y_dev || !out_dev || phonemes <= 0 || phonemes > kMaxPhonemes || frames < 0 ||
        x_stride < frames || y_stride < frames) {
        set_error("ppg_distance: bad argument (phonemes must be in [1, %d])", kMaxPhonemes);
        return PPGS_E_INVALID;
    }
    if (reduction < 0 || reduction > 2) {
        set_error("Reduction method %d not defined", reduction);
        return PPGS_E_INVALID;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int P = phonemes;
    const size_t weights_bytes = ((size_t)P * P * 4 + 255) & ~size_t(255);
    PPGS_CHECK(ensure_workspace(e, weights_bytes + (size_t)(frames > 0 ? frames : 1) * 4));
    e->cached_plan_dev = nullptr;   // the scratch below overwrites a model engine's resident plan tables
    float* weights = static_cast<float*>(e->workspace);
    float* per_frame = reduction == 0 ? out_dev : reinterpret_cast<float*>(static_cast<char*>(e->workspace) + weights_bytes);
    if (similarity_dev) {
        LaunchScope scope(e, "ppg_similarity_weights", stream);
        similarity_weights_kernel<<<P, 64, 0, stream>>>(similarity_dev, P, exponent, weights);
    }
    if (frames > 0) {
        const int warps = 8;
        const size_t smem = ((similarity_dev ? (size_t)P * P : 0) + (size_t)warps * 2 * P) * 4;
        LaunchScope scope(e, "ppg_distance", stream);
        distance_kernel<<<(unsigned)((frames + warps - 1) / warps), warps * 32, smem, stream>>>(
            x_dev, y_dev, P, frames, x_stride, y_stride, similarity_dev ? weights : nullptr, per_frame);
    }
    PPGS_CUDA(cudaGetLastError());
    if (reduction != 0) {
        LaunchScope scope(e, "ppg_reduce", stream);
        const float scale = reduction == 1 ? 1.f / (float)frames : 1.f;   // mean of nothing = nan, like torch
        reduce_sum_kernel<<<1, 1024, 0, stream>>>(per_frame, frames, scale, out_dev);
        PPGS_CUDA(cudaGetLastError());
    }
    return PPGS_OK;
}

int ppgs_ppg_interpolate(ppgs_engine* e, const float* x_dev, const float* y_dev, const float* interp_dev,
                         float interp_scalar, int64_t rows, int64_t frames, float* out_dev, void* stream_) {
    PPGS_POST_ENTER(e);
    if (!x_dev || !y_dev || !out_dev || rows < 0 || frames < 0) {
        set_error("ppg_interpolate: bad argument");
        return PPGS_E_INVALID;
    }
    const int64_t count = rows * frames;
    if (count == 0) return PPGS_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    {
        LaunchScope scope(e, "ppg_interpolate", stream);
        interpolate_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(x_dev, y_dev, interp_dev,
                                                                              interp_scalar, frames, count, out_dev);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

int ppgs_ppg_grid_sample(ppgs_engine* e, const float* ppg_dev, int phonemes, int64_t frames, const float* grid_dev,
                         int64_t grid_len, float* out_dev, void* stream_) {
    PPGS_POST_ENTER(e);
    if (!ppg_dev || !grid_dev || !out_dev || phonemes <= 0 || frames <= 0 || grid_len < 0) {
        set_error("ppg_grid_sample: bad argument");
        return PPGS_E_INVALID;
    }
    if (grid_len == 0) return PPGS_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    {
        LaunchScope scope(e, "ppg_grid_sample", stream);
        grid_sample_kernel<<<(unsigned)((grid_len + 127) / 128), 128, 0, stream>>>(ppg_dev, phonemes, frames, grid_dev,
                                                                                 grid_len, out_dev);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

int ppgs_ppg_sparsify(ppgs_engine* e, const float* ppg_dev, int batch, int phonemes, int64_t frames, int method,
                      float threshold, float* out_dev, void* stream_) {
    PPGS_POST_ENTER(e);
    if (!ppg_dev || !out_dev || batch < 0 || phonemes <= 0 || phonemes > kMaxPhonemes || frames < 0) {
        set_error("ppg_sparsify: bad argument (phonemes must be in [1, %d])", kMaxPhonemes);
        return PPGS_E_INVALID;
    }
    if (method < kSparsifyConstant || method > kSparsifyTopk) {
        set_error("Sparsification method %d not defined", method);
        return PPGS_E_INVALID;
    }
    if (method == kSparsifyPercentile && !(threshold >= 0.f && threshold <= 1.f)) {
        set_error("quantile() q values must be in the range [0, 1]");
        return PPGS_E_INVALID;
    }
    if (method == kSparsifyTopk && (threshold < 1.f || threshold > (float)phonemes)) {
        set_error("selected index k out of range");
        return PPGS_E_INVALID;
    }
    const int64_t total = (int64_t)batch * frames;
    if (total == 0) return PPGS_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    {
        LaunchScope scope(e, "ppg_sparsify", stream);
        sparsify_kernel<<<(unsigned)((total + 7) / 8), 256, 0, stream>>>(ppg_dev, phonemes, frames, total, method,
                                                                       threshold, out_dev);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

}  // extern "C"
