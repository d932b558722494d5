// Validation entry points (debug_abi.h): single-kernel runs on host data.
#include "debug_abi.h"

#include <vector>

#include "attention_tc.cuh"
#include "gemm_tc.cuh"

using namespace ppgs;

namespace {

struct DeviceBuffer {
    void* ptr = nullptr;
    ~DeviceBuffer() { cudaFree(ptr); }
    int alloc(size_t bytes) {
        PPGS_CUDA(cudaMalloc(&ptr, bytes));
        PPGS_CUDA(cudaMemset(ptr, 0, bytes));
        return PPGS_OK;
    }
    template <typename T>
    T* as() { return static_cast<T*>(ptr); }
};

void split_planes(const float* src, size_t count, std::vector<__half>& planes) {
    planes.resize(2 * count);
    for (size_t i = 0; i < count; ++i) {
        const __half hi = __float2half_rn(src[i]);
        planes[i] = hi;
        planes[count + i] = __float2half_rn(src[i] - __half2float(hi));
    }
}

int ensure_status(ppgs_engine* e) {
    if (!e->status_dev) {
        PPGS_CUDA(cudaMalloc(&e->status_dev, sizeof(int)));
        PPGS_CUDA(cudaMemset(e->status_dev, 0, sizeof(int)));
    }
    return PPGS_OK;
}

}  // namespace

extern "C" int ppgs_debug_gemm(ppgs_engine* e, const float* a_host, const float* w_host,
                               const float* bias_host, int M, int N, int C, int taps, int bn,
                               int pair, int a_planes, int b_planes, float* out_host) {
    if (!e || !a_host || !w_host || !bias_host || !out_host || M <= 0 || M % 128 || C % 8 ||
        N <= 0 || taps <= 0 || (taps & 1) == 0 || (pair && (M % 256 || bn != 256))) {
        set_error("debug_gemm: bad argument");
        return PPGS_E_INVALID;
    }
    PPGS_CUDA(cudaSetDevice(e->device));
    PPGS_CHECK(ensure_status(e));
    std::vector<__half> a_planes_host, w_planes_host;
    split_planes(a_host, (size_t)M * C, a_planes_host);
    HostTensor w;
    w.data.assign(w_host, w_host + (size_t)N * C * taps);
    w.shape = {N, C, taps};
    const float inv_scale = pack_planes(&w, w_planes_host);

    DeviceBuffer a_dev, w_dev, bias_dev, scale_dev, out_dev;
    PPGS_CHECK(a_dev.alloc(a_planes_host.size() * 2));
    PPGS_CHECK(w_dev.alloc(w_planes_host.size() * 2));
    PPGS_CHECK(bias_dev.alloc((size_t)N * 4));
    PPGS_CHECK(scale_dev.alloc(4));
    PPGS_CHECK(out_dev.alloc((size_t)M * N * 4));
    PPGS_CUDA(cudaMemcpy(a_dev.ptr, a_planes_host.data(), a_planes_host.size() * 2, cudaMemcpyHostToDevice));
    PPGS_CUDA(cudaMemcpy(w_dev.ptr, w_planes_host.data(), w_planes_host.size() * 2, cudaMemcpyHostToDevice));
    PPGS_CUDA(cudaMemcpy(bias_dev.ptr, bias_host, (size_t)N * 4, cudaMemcpyHostToDevice));
    PPGS_CUDA(cudaMemcpy(scale_dev.ptr, &inv_scale, 4, cudaMemcpyHostToDevice));

    CUtensorMap map_a, map_b;
    PPGS_CHECK(tc::make_plane_map(&map_a, a_dev.as<__half>(), false, C, M, 1, 2, C, 0,
                                  (uint64_t)M * C, 128, a_planes));
    PPGS_CHECK(tc::make_plane_map(&map_b, w_dev.as<__half>(), true, C, N, taps, 2, C,
                                  (uint64_t)N * C, (uint64_t)taps * N * C, pair ? bn / 2 : bn, b_planes));
    tc::GemmParams p;
    p.m_tiles = M / 128;
    p.n_tiles = (N + bn - 1) / bn;
    p.taps = taps;
    p.half = taps / 2;
    p.cblocks = (C + 63) / 64;
    p.a_planes = a_planes;
    p.b_planes = b_planes;
    p.pair = pair;
    p.N = N;
    p.scale = scale_dev.as<float>();
    p.bias = bias_dev.as<float>();
    p.out_f32 = out_dev.as<float>();
    p.ld_f32 = N;
    p.status = e->status_dev;
    PPGS_CHECK(tc::launch_gemm_tc(e, "debug_gemm", bn, tc::kEpiF32, map_a, map_b, nullptr, p, nullptr));
    PPGS_CUDA(cudaDeviceSynchronize());
    PPGS_CHECK(check_status(e, nullptr));
    PPGS_CUDA(cudaMemcpy(out_host, out_dev.ptr, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
    return PPGS_OK;
}

extern "C" int ppgs_debug_attention(ppgs_engine* e, const float* qkv_host, int rows, int tensor_len,
                                    int valid_len, int H, int heads, int causal, int planes,
                                    int use_tensor_cores, float* out_host) {
    if (!e || !qkv_host || !out_host || rows <= 0 || rows % 128 || tensor_len > rows ||
        valid_len > tensor_len || heads <= 0 || H % heads) {
        set_error("debug_attention: bad argument");
        return PPGS_E_INVALID;
    }
    PPGS_CUDA(cudaSetDevice(e->device));
    PPGS_CHECK(ensure_status(e));
    std::vector<__half> qkv_planes;
    split_planes(qkv_host, (size_t)rows * 3 * H, qkv_planes);
    SeqInfo s{};
    s.row0 = 0;
    s.tensor_len = tensor_len;
    s.valid_len = valid_len;
    s.keep_end = tensor_len;
    DeviceBuffer qkv_dev, out_dev, seq_dev;
    PPGS_CHECK(qkv_dev.alloc(qkv_planes.size() * 2));
    PPGS_CHECK(out_dev.alloc((size_t)2 * rows * H * 2));
    PPGS_CHECK(seq_dev.alloc(sizeof(SeqInfo)));
    PPGS_CUDA(cudaMemcpy(qkv_dev.ptr, qkv_planes.data(), qkv_planes.size() * 2, cudaMemcpyHostToDevice));
    PPGS_CUDA(cudaMemcpy(seq_dev.ptr, &s, sizeof(SeqInfo), cudaMemcpyHostToDevice));
    const int saved = e->attention_impl;
    e->attention_impl = use_tensor_cores ? 1 : 0;
    // use_tensor_cores == 2: Q / K / P as single fp16 planes (head_dim 128: the two-tile kernel)
    const int operand_planes = use_tensor_cores == 2 ? 1 : 2;
    const int rc = launch_attention_any(e, qkv_dev.as<__half>(), out_dev.as<__half>(), rows, H, heads, rows, 1,
                                        seq_dev.as<SeqInfo>(), causal, planes, nullptr, 0, 0, operand_planes,
                                        operand_planes);
    e->attention_impl = saved;
    PPGS_CHECK(rc);
    PPGS_CUDA(cudaDeviceSynchronize());
    PPGS_CHECK(check_status(e, nullptr));
    std::vector<__half> out_planes((size_t)2 * rows * H);
    PPGS_CUDA(cudaMemcpy(out_planes.data(), out_dev.ptr, out_planes.size() * 2, cudaMemcpyDeviceToHost));
    const size_t plane = (size_t)rows * H;
    for (size_t i = 0; i < plane; ++i)
        out_host[i] = __half2float(out_planes[i]) + __half2float(out_planes[plane + i]);
    return PPGS_OK;
}

extern "C" int ppgs_debug_trace(ppgs_engine* e, unsigned long long* out64) {
    if (!e || !out64) {
        set_error("debug_trace: NULL argument");
        return PPGS_E_INVALID;
    }
    if (!e->trace_dev) {
        set_error("debug_trace: create the engine with PPGS_B200_TRACE=1");
        return PPGS_E_STATE;
    }
    PPGS_CUDA(cudaSetDevice(e->device));
    PPGS_CUDA(cudaDeviceSynchronize());
    PPGS_CUDA(cudaMemcpy(out64, e->trace_dev, 128 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    PPGS_CUDA(cudaMemset(e->trace_dev, 0, 128 * sizeof(unsigned long long)));
    return PPGS_OK;
}
