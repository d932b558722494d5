// Split-fp16 tensor-core GEMM for sm_100a: C[M,N] = A[M,K] * W[N,K]^T with both
// operands stored as (hi, lo) fp16 planes, products accumulated in fp32 in TMEM:
//   acc = A_hi*W_hi + A_hi*W_lo + A_lo*W_hi          (PPGS_PRECISION_F16X2)
//   acc = A_hi*W_hi                                  (PPGS_PRECISION_F16)
// It carries every dense contraction of ppgs/model/transformer.py:65-81 (input /
// output Conv1d as `taps` shifted GEMMs, MultiheadAttention in/out projections,
// linear1/linear2 of torch.nn.TransformerEncoderLayer) with the surrounding
// elementwise work fused into the TMEM epilogue.
//
// Structure (persistent, one CTA per SM, 192 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor boxes {64 x 128 x planes} of A and
//              {64 x BN x planes} of W into a 128B-swizzled shared-memory ring
//   warp 1     tcgen05.mma issuer (one elected lane), M=128 x N=BN x K=16 per
//              instruction, accumulator double-buffered in TMEM (2 x BN columns)
//   warps 2-5  epilogue: tcgen05.ld 32 lanes x 32 columns, one output row per thread
#include "gemm_tc.cuh"

namespace ppgs {
namespace tc {

constexpr int kGemmThreads = 192;
constexpr int kPlaneABytes = kBM * kBK * 2;   // 16 KB

template <int BN>
struct GemmShape {
    static constexpr int kPlaneBBytes = BN * kBK * 2;
    static constexpr int kStageBytes = 2 * kPlaneABytes + 2 * kPlaneBBytes;
    static constexpr int kStages = (BN == 256) ? 2 : (BN == 128) ? 3 : 4;
    static constexpr int kAccCols = BN;                      // power of two >= 32
    static constexpr int kTmemCols = 2 * kAccCols;
    static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024;
};

__device__ __forceinline__ void store_f32x32(float* dst, const float (&y)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(dst)[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
}

__device__ __forceinline__ void store_planes32(__half* hi_dst, __half* lo_dst, const float (&y)[32]) {
    uint32_t h[16], l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        __half h0, l0, h1, l1;
        split_f16(y[2 * i], h0, l0);
        split_f16(y[2 * i + 1], h1, l1);
        h[i] = pack_half2(h0, h1);
        l[i] = pack_half2(l0, l1);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        reinterpret_cast<uint4*>(hi_dst)[i] = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
        reinterpret_cast<uint4*>(lo_dst)[i] = make_uint4(l[4 * i], l[4 * i + 1], l[4 * i + 2], l[4 * i + 3]);
    }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const GemmParams p) {
    using Shape = GemmShape<BN>;
    constexpr int kStages = Shape::kStages;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ __align__(8) uint64_t empty_bar[kStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles;
    const int num_kb = p.taps * p.cblocks;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Shape::kTmemCols>(&tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            prefetch_tensormap(&map_a);
            prefetch_tensormap(&map_b);
            const uint32_t stage_tx = p.a_planes * kPlaneABytes + p.b_planes * Shape::kPlaneBBytes;
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
                const int m_blk = tile / p.n_tiles, n_blk = tile - m_blk * p.n_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
                    if (!mbar_wait(&empty_bar[stage], phase ^ 1)) {
                        atomicExch(p.status, kStatusProducerTimeout);
                        ok = false;
                        break;
                    }
                    unsigned char* sa = smem + (size_t)stage * Shape::kStageBytes;
                    unsigned char* sb = sa + 2 * kPlaneABytes;
                    mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                    tma_load_3d(sa, &map_a, &full_bar[stage], cb * kBK, m_blk * kBM + tap - p.half, 0);
                    tma_load_4d(sb, &map_b, &full_bar[stage], cb * kBK, n_blk * BN, tap, 0);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(kBM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x, ++iter) {
                const int acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                if (!mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1)) {
                    atomicExch(p.status, kStatusMmaTimeout);
                    break;
                }
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Shape::kAccCols;
                uint32_t accumulate = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (!mbar_wait(&full_bar[stage], phase)) {
                        atomicExch(p.status, kStatusMmaTimeout);
                        ok = false;
                        break;
                    }
                    tcgen05_fence_after();
                    const uint32_t a0 = smem_u32(smem + (size_t)stage * Shape::kStageBytes);
                    const uint32_t a1 = a0 + kPlaneABytes;
                    const uint32_t b0 = a0 + 2 * kPlaneABytes;
                    const uint32_t b1 = b0 + Shape::kPlaneBBytes;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        const uint32_t koff = k * 32;   // 16 fp16 along K inside the swizzle row
                        const uint64_t da0 = smem_desc_kmajor_sw128(a0 + koff);
                        const uint64_t db0 = smem_desc_kmajor_sw128(b0 + koff);
                        umma_f16(d_tmem, da0, db0, idesc, accumulate);
                        accumulate = 1;
                        if (p.b_planes == 2)
                            umma_f16(d_tmem, da0, smem_desc_kmajor_sw128(b1 + koff), idesc, 1);
                        if (p.a_planes == 2)
                            umma_f16(d_tmem, smem_desc_kmajor_sw128(a1 + koff), db0, idesc, 1);
                    }
                    umma_commit(&empty_bar[stage]);   // frees the smem slot when the MMAs retire
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (ok) umma_commit(&tmem_full_bar[acc]);
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue
        const int quad = warp & 3;                  // TMEM lane quadrant of this warp
        const int row_in_tile = quad * 32 + lane;
        const float scale = *p.scale;
        int iter = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
            const int m_blk = tile / p.n_tiles, n_blk = tile - m_blk * p.n_tiles;
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            if (!mbar_wait(&tmem_full_bar[acc], acc_phase)) {
                atomicExch(p.status, kStatusEpilogueTimeout);
                break;
            }
            tcgen05_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * Shape::kAccCols;
            const int64_t m = (int64_t)m_blk * kBM + row_in_tile;
            uint32_t raw[32];
            float y[32];

            if (EPI == kEpiF32 || EPI == kEpiPlanes) {
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    tmem_ld_32x32(t_acc + c * 32, raw);
                    tmem_wait_ld();
                    const int n0 = n_blk * BN + c * 32;
                    if (n0 >= p.N) continue;   // warp-uniform
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float v = __uint_as_float(raw[j]) * scale;
                        if (n0 + j < p.N) v += __ldg(p.bias + n0 + j);
                        if (EPI == kEpiPlanes && p.relu) v = fmaxf(v, 0.f);
                        y[j] = v;
                    }
                    if (EPI == kEpiF32) {
                        float* dst = p.out_f32 + m * p.ld_f32 + n0;
                        if (n0 + 32 <= p.N && (p.ld_f32 & 3) == 0) store_f32x32(dst, y);
                        else
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < p.N) dst[j] = y[j];
                    } else {
                        __half* dst = p.out_planes + m * p.ld_planes + n0;
                        store_planes32(dst, dst + p.plane_stride, y);
                    }
                }
            } else if (EPI == kEpiConvIn) {
                const SeqInfo s = p.seqs[p.tile_seq[m_blk]];
                const int t = (int)(m - s.row0);
                const bool in_tensor = t < s.tensor_len, valid = t < s.valid_len;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    tmem_ld_32x32(t_acc + c * 32, raw);
                    tmem_wait_ld();
                    const int n0 = n_blk * BN + c * 32;
                    const float* pe = p.pe + (int64_t)(in_tensor ? t : 0) * p.N + n0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float v = valid ? __uint_as_float(raw[j]) * scale + __ldg(p.bias + n0 + j) : 0.f;
                        y[j] = in_tensor ? v + pe[j] : 0.f;
                    }
                    store_f32x32(p.out_f32 + m * p.ld_f32 + n0, y);
                    __half* dst = p.out_planes + m * p.ld_planes + n0;
                    store_planes32(dst, dst + p.plane_stride, y);
                }
            } else if (EPI == kEpiResLN) {
                // the tile is one full row of the hidden state (BN == N == hidden)
                const SeqInfo s = p.seqs[p.tile_seq[m_blk]];
                const bool in_tensor = (int)(m - s.row0) < s.tensor_len;
                const float* res = p.residual + m * p.N;
                float sum = 0.f;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    tmem_ld_32x32(t_acc + c * 32, raw);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 r = reinterpret_cast<const float4*>(res + c * 32)[j];
                        const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int n = c * 32 + 4 * j + q;
                            const float v = __uint_as_float(raw[4 * j + q]) * scale + __ldg(p.bias + n) + rr[q];
                            sum += v;
                            raw[4 * j + q] = __float_as_uint(v);
                        }
                    }
                    tmem_st_32x32(t_acc + c * 32, raw);
                }
                tmem_wait_st();
                const float mean = sum * (1.f / BN);
                float sq = 0.f;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    tmem_ld_32x32(t_acc + c * 32, raw);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float d = __uint_as_float(raw[j]) - mean;
                        sq = fmaf(d, d, sq);
                    }
                }
                const float rstd = rsqrtf(sq * (1.f / BN) + p.eps);
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    tmem_ld_32x32(t_acc + c * 32, raw);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = c * 32 + j;
                        const float v = (__uint_as_float(raw[j]) - mean) * rstd * __ldg(p.gamma + n) +
                                        __ldg(p.beta + n);
                        y[j] = in_tensor ? v : 0.f;
                    }
                    store_f32x32(p.out_f32 + m * p.ld_f32 + c * 32, y);
                    __half* dst = p.out_planes + m * p.ld_planes + c * 32;
                    store_planes32(dst, dst + p.plane_stride, y);
                }
            } else if (EPI == kEpiConvOut) {
                // BN == 64 >= O: all channels of a frame live in this thread
                const SeqInfo s = p.seqs[p.tile_seq[m_blk]];
                const int t = (int)(m - s.row0);
                const bool valid = t < s.valid_len;
                float z[64];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tmem_ld_32x32(t_acc + c * 32, raw);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = c * 32 + j;
                        z[n] = (valid && n < p.O) ? __uint_as_float(raw[j]) * scale + __ldg(p.bias + n) : 0.f;
                    }
                }
                if (t >= s.keep_begin && t < s.keep_end) {
                    float inv = 1.f, mx = 0.f;
                    if (p.softmax) {
                        mx = -3.0e38f;
#pragma unroll
                        for (int n = 0; n < 64; ++n)
                            if (n < p.O) mx = fmaxf(mx, z[n]);
                        float sum = 0.f;
#pragma unroll
                        for (int n = 0; n < 64; ++n)
                            if (n < p.O) {
                                z[n] = expf(z[n] - mx);
                                sum += z[n];
                            }
                        inv = 1.f / sum;
                    }
                    float* dst = p.ppg + (int64_t)s.batch * p.O * p.T + s.out_start + (t - s.keep_begin);
#pragma unroll
                    for (int n = 0; n < 64; ++n)
                        if (n < p.O) dst[(int64_t)n * p.T] = z[n] * inv;
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<Shape::kTmemCols>(tmem_base);
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult query;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &query) ==
                cudaSuccess &&
            query == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

int make_plane_map(CUtensorMap* map, const __half* base, bool rank4, uint64_t inner,
                   uint64_t rows, uint64_t groups, uint64_t planes, uint64_t row_stride_elems,
                   uint64_t group_stride_elems, uint64_t plane_stride_elems, uint32_t box_rows,
                   uint32_t box_planes) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return PPGS_E_CUDA;
    }
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4], elem[4] = {1, 1, 1, 1};
    int rank;
    if (!rank4) {
        if (groups != 1) {
            set_error("make_plane_map: rank-3 maps have one group");
            return PPGS_E_INVALID;
        }
        rank = 3;
        dims[0] = inner; dims[1] = rows; dims[2] = planes;
        strides[0] = row_stride_elems * 2; strides[1] = plane_stride_elems * 2;
        box[0] = kBK; box[1] = box_rows; box[2] = box_planes;
    } else {
        rank = 4;
        dims[0] = inner; dims[1] = rows; dims[2] = groups; dims[3] = planes;
        strides[0] = row_stride_elems * 2; strides[1] = group_stride_elems * 2;
        strides[2] = plane_stride_elems * 2;
        box[0] = kBK; box[1] = box_rows; box[2] = 1; box[3] = box_planes;
    }
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<__half*>(base), dims,
                     strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner %llu rows %llu groups %llu "
                  "planes %llu)", (int)rc, (unsigned long long)inner, (unsigned long long)rows,
                  (unsigned long long)groups, (unsigned long long)planes);
        return PPGS_E_CUDA;
    }
    return PPGS_OK;
}

template <int BN, int EPI>
static int launch_one(ppgs_engine* e, const char* name, const CUtensorMap& map_a,
                      const CUtensorMap& map_b, const GemmParams& p, cudaStream_t stream) {
    using Shape = GemmShape<BN>;
    static bool attr = false;
    if (!attr) {
        PPGS_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Shape::kSmemBytes));
        attr = true;
    }
    const int tiles = p.m_tiles * p.n_tiles;
    const int grid = std::min(tiles, e->sm_count);
    {
        LaunchScope scope(e, name, stream);
        gemm_tc_kernel<BN, EPI><<<grid, kGemmThreads, Shape::kSmemBytes, stream>>>(map_a, map_b, p);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

int launch_gemm_tc(ppgs_engine* e, const char* name, int bn, int epilogue,
                   const CUtensorMap& map_a, const CUtensorMap& map_b, const GemmParams& p,
                   cudaStream_t stream) {
    if (p.m_tiles <= 0 || p.n_tiles <= 0 || p.cblocks <= 0 || !p.scale || !p.status) {
        set_error("gemm_tc: bad parameters");
        return PPGS_E_INVALID;
    }
#define PPGS_GEMM_CASE(BN_, EPI_) \
    if (bn == BN_ && epilogue == EPI_) return launch_one<BN_, EPI_>(e, name, map_a, map_b, p, stream)
    PPGS_GEMM_CASE(256, kEpiF32);
    PPGS_GEMM_CASE(128, kEpiF32);
    PPGS_GEMM_CASE(64, kEpiF32);
    PPGS_GEMM_CASE(256, kEpiPlanes);
    PPGS_GEMM_CASE(256, kEpiConvIn);
    PPGS_GEMM_CASE(256, kEpiResLN);
    PPGS_GEMM_CASE(64, kEpiConvOut);
#undef PPGS_GEMM_CASE
    set_error("gemm_tc: no kernel for BN=%d epilogue=%d", bn, epilogue);
    return PPGS_E_UNSUPPORTED;
}

}  // namespace tc
}  // namespace ppgs
