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

constexpr int kGemmThreads = 320;      // producer + MMA + 8 epilogue warps
constexpr int kPlaneABytes = kBM * kBK * 2;   // 16 KB
constexpr int kStageOutBytes = kBM * 64 * 2;   // [128 rows][64 cols] fp16 = 16 KB, one plane

// PAIR: two CTAs of a cluster form one cta_group::2 MMA (M = 256): each loads its
// own 128 rows of A and HALF of the W tile, so a stage is 64 KB instead of 96 KB
// per SM at BN = 256 and three stages fit.
template <int BN, bool PAIR>
struct GemmShape {
    static constexpr int kBRows = PAIR ? BN / 2 : BN;          // W rows held by this CTA
    static constexpr int kPlaneBBytes = kBRows * kBK * 2;
    static constexpr int kStageBytes = 2 * kPlaneABytes + 2 * kPlaneBBytes;
    static constexpr int kStages = PAIR ? 3 : (BN == 256) ? 2 : (BN == 128) ? 3 : 4;
    static constexpr int kAccCols = BN;                      // power of two >= 32
    static constexpr int kTmemCols = 2 * kAccCols;
    static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 2 * kStageOutBytes;
};

// One plane (hi or lo) of one 64-column group of one output row -> this thread's
// 128-byte row of the staging tile ([128 rows][64 cols] fp16, 128-byte swizzle), then
// one elected thread of the 128-thread epilogue set stores the tile with TMA.  Rows
// of 128 contiguous bytes keep the number of L2 write requests per byte at its
// minimum (64-byte rows cost twice the requests and made the stores the bottleneck).
__device__ __forceinline__ void stage_store_plane(const CUtensorMap* map_out, unsigned char* stage,
                                                  int set, int row, bool elected,
                                                  const uint32_t (&w)[32], int n0, int m0, int plane,
                                                  int debug_flags) {
    if (debug_flags & 1) return;
    if (elected && !(debug_flags & 8)) bulk_wait_read_all();   // previous store has drained the buffer
    named_bar_sync(1 + set, 128);
    const uint32_t base = smem_u32(stage) + (uint32_t)row * 128;
    const uint32_t sw = (uint32_t)row & 7;
#pragma unroll
    for (int u = 0; u < 8; ++u)
        st_shared_v4(base + (((uint32_t)u ^ sw) << 4), w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
    fence_proxy_async_smem();
    named_bar_sync(1 + set, 128);
    if (elected && !(debug_flags & 4)) {
        tma_store_3d(map_out, stage, n0, m0, plane);
        bulk_commit_group();
    }
}

// 32 consecutive fp32 parameters (bias / gamma / beta / PE row), same address in
// every lane: 8 broadcast 128-bit loads through the read-only path
__device__ __forceinline__ void load_params32(const float* src, float (&out)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + j);
        out[4 * j] = v.x;
        out[4 * j + 1] = v.y;
        out[4 * j + 2] = v.z;
        out[4 * j + 3] = v.w;
    }
}

// 32 values -> 16 packed (hi, lo) words at h[at .. at + 16), l[at .. at + 16)
__device__ __forceinline__ void split32(const float (&y)[32], uint32_t (&h)[32], uint32_t (&l)[32],
                                        int at) {
#pragma unroll
    for (int i = 0; i < 16; ++i) split2_f16(y[2 * i], y[2 * i + 1], h[at + i], l[at + i]);
}

// work item (row tile, or tile pair in the CTA-pair kernel) -> row tile of this CTA; identity
// unless a row-tile window is set (GemmParams::win_*, in tiles).  A pair is two ADJACENT tiles;
// it need not start at an even tile.
template <bool PAIR>
__device__ __forceinline__ int window_tile(const GemmParams& p, int item, int rank) {
    if (p.reverse) item = (PAIR ? p.m_tiles / 2 : p.m_tiles) - 1 - item;
    if (p.win_size == 0) return PAIR ? 2 * item + rank : item;
    const int per_seq = PAIR ? p.win_size / 2 : p.win_size;
    const int seq = item / per_seq, u = item - seq * per_seq;
    const int first = p.win_per_seq ? p.seqs[seq].src_start : p.win_first;
    return seq * p.win_stride + first + (PAIR ? 2 * u + rank : u);
}

template <int BN, int EPI, bool PAIR>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_out, const GemmParams p) {
    using Shape = GemmShape<BN, PAIR>;
    constexpr int kStages = Shape::kStages;
    constexpr int kChunks = BN / 32;
    // 128B-swizzled TMA / UMMA tiles need 1024-byte alignment: the declaration
    // aligns the dynamic region (checked below), there is no room for slack at BN=256
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* stage_out = smem + (size_t)kStages * Shape::kStageBytes;
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ __align__(8) uint64_t empty_bar[kStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float ln_part[2][kBM];      // [epilogue set][row] partial row statistics

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work items: (m tile, n tile) per CTA, or (pair of m tiles, n tile) per CTA pair
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int total_tiles = (PAIR ? p.m_tiles / 2 : p.m_tiles) * p.n_tiles;
    const int num_kb = p.taps * p.cblocks;
    if (smem_u32(smem) & 1023u) {   // uniform across the CTA
        if (threadIdx.x == 0) atomicExch(p.status, kStatusBadAlignment);
        return;
    }

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], PAIR ? 16 : 8);   // leader: epilogue warps of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (PAIR) tmem_alloc_pair<Shape::kTmemCols>(&tmem_slot);
        else tmem_alloc<Shape::kTmemCols>(&tmem_slot);
    }
    tcgen05_fence_before();
    if (PAIR) cluster_sync_all();   // peer barriers are initialised before any remote signal
    else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;
    // the set-up above may overlap the previous kernel; its results are needed from here on
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            prefetch_tensormap(&map_a);
            prefetch_tensormap(&map_b);
            const uint32_t stage_tx = p.a_planes * kPlaneABytes + p.b_planes * Shape::kPlaneBBytes;
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            long long t_wait = 0;
            const long long t_begin = clock64();
            for (int tile = worker; tile < total_tiles && ok; tile += workers) {
                const int m_item = tile / p.n_tiles, n_blk = tile - m_item * p.n_tiles;
                const int m_blk = window_tile<PAIR>(p, m_item, (int)rank);
                const int a_row0 = m_blk * kBM * p.row_mul - p.half + n_blk * p.a_group_rows;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks + p.cb0;
                    const long long t0 = clock64();
                    const bool got = mbar_wait(&empty_bar[stage], phase ^ 1);
                    t_wait += clock64() - t0;
                    if (!got) {
                        atomicExch(p.status, kStatusProducerTimeout);
                        ok = false;
                        break;
                    }
                    unsigned char* sa = smem + (size_t)stage * Shape::kStageBytes;
                    unsigned char* sb = sa + 2 * kPlaneABytes;
                    if (PAIR) {
                        // the leader's barrier collects the bytes of both CTAs
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_tx);
                        const uint32_t leader_bar = map_to_cta(&full_bar[stage], 0);
                        tma_load_3d_pair(sa, &map_a, leader_bar, cb * kBK, a_row0 + tap, 0);
                        tma_load_4d_pair(sb, &map_b, leader_bar, cb * kBK,
                                         n_blk * BN + (int)rank * Shape::kBRows, tap, 0);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                        tma_load_3d(sa, &map_a, &full_bar[stage], cb * kBK, a_row0 + tap, 0);
                        tma_load_4d(sb, &map_b, &full_bar[stage], cb * kBK, n_blk * BN, tap, 0);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            if (p.trace) {
                atomicAdd(p.trace + 5, (unsigned long long)t_wait);
                atomicAdd(p.trace + 6, (unsigned long long)(clock64() - t_begin));
                atomicAdd(p.trace + 7, 1ull);
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------- MMA issuer
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc_f16(PAIR ? 2 * kBM : kBM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            bool ok = true;
            long long t_empty = 0, t_full = 0;
            const long long t_begin = clock64();
            for (int tile = worker; tile < total_tiles && ok; tile += workers, ++iter) {
                // split_acc: both accumulator buffers belong to ONE tile (hi.hi products in the
                // first, the two correction products in the second; the epilogue adds them)
                const int acc = p.split_acc ? 0 : iter & 1;
                const uint32_t acc_phase = p.split_acc ? (iter & 1) : (iter >> 1) & 1;
                const long long t0 = clock64();
                const bool got = mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                t_empty += clock64() - t0;
                if (!got) {
                    atomicExch(p.status, kStatusMmaTimeout);
                    break;
                }
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Shape::kAccCols;
                const uint32_t d_corr = p.split_acc ? tmem_base + Shape::kAccCols : d_tmem;
                uint32_t accumulate = 0, accumulate_corr = p.split_acc ? 0u : 1u;
                // N blocks whose outputs are kept as one fp16 plane may run fewer passes
                const int n_blk_mma = tile - (tile / p.n_tiles) * p.n_tiles;
                const int passes = (n_blk_mma + 1) * BN <= p.hi_only_cols && !p.split_acc ? p.hi_only_passes : 3;
                const bool pass_b_lo = p.b_planes == 2 && passes >= 3;
                const bool pass_a_lo = p.a_planes == 2 && passes >= 2;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const long long t1 = clock64();
                    const bool landed = mbar_wait(&full_bar[stage], phase);
                    t_full += clock64() - t1;
                    if (!landed) {
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
                        if (PAIR) {
                            umma_f16_pair(d_tmem, da0, db0, idesc, accumulate);
                            if (pass_b_lo) {
                                umma_f16_pair(d_corr, da0, smem_desc_kmajor_sw128(b1 + koff), idesc, accumulate_corr);
                                accumulate_corr = 1;
                            }
                            if (pass_a_lo) {
                                umma_f16_pair(d_corr, smem_desc_kmajor_sw128(a1 + koff), db0, idesc, accumulate_corr);
                                accumulate_corr = 1;
                            }
                        } else {
                            umma_f16(d_tmem, da0, db0, idesc, accumulate);
                            if (pass_b_lo)
                                umma_f16(d_tmem, da0, smem_desc_kmajor_sw128(b1 + koff), idesc, 1);
                            if (pass_a_lo)
                                umma_f16(d_tmem, smem_desc_kmajor_sw128(a1 + koff), db0, idesc, 1);
                        }
                        accumulate = 1;
                    }
                    // frees the smem slot (in both CTAs of a pair) when the MMAs retire
                    if (PAIR) umma_commit_pair(&empty_bar[stage]);
                    else umma_commit(&empty_bar[stage]);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (ok) {
                    if (PAIR) umma_commit_pair(&tmem_full_bar[acc]);
                    else umma_commit(&tmem_full_bar[acc]);
                }
            }
            if (p.trace) {
                atomicAdd(p.trace + 0, (unsigned long long)t_empty);
                atomicAdd(p.trace + 1, (unsigned long long)t_full);
                atomicAdd(p.trace + 2, (unsigned long long)(clock64() - t_begin));
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue
        // 8 warps = 2 sets of 4; a set covers the 128 rows (TMEM lane quadrant =
        // warp % 4) and takes every other 32-column chunk.
        const int set = (warp - 2) >> 2;
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const bool elected = (warp - 2) == 4 * set && lane == 0;
        const int tid_in_set = (int)threadIdx.x - 64 - 128 * set;
        unsigned char* stage = stage_out + set * kStageOutBytes;
        const float scale = *p.scale;
        int iter = 0;
        long long t_wait = 0;
        const long long t_begin = clock64();
        for (int tile = worker; tile < total_tiles; tile += workers, ++iter) {
            const int m_item = tile / p.n_tiles, n_blk = tile - m_item * p.n_tiles;
            const int m_blk = window_tile<PAIR>(p, m_item, (int)rank);
            const int acc = p.split_acc ? 0 : iter & 1;
            const uint32_t acc_phase = p.split_acc ? (iter & 1) : (iter >> 1) & 1;
            // ResLN: the residual loads of the first two column chunks are in flight while
            // this warp waits for the accumulator
            uint4 pre[2][8];
            auto coop_load = [&](int c, uint4 (&dst)[8]) {
#pragma unroll
                for (int it8 = 0; it8 < 8; ++it8) {
                    const int idx = it8 * 128 + tid_in_set;
                    const int u = idx & 3, r = (idx >> 2) & 127, plane = idx >> 9;
                    dst[it8] = ld_global_stream(p.residual + plane * p.res_plane_stride +
                                                (int64_t)(m_blk * kBM + r) * p.res_ld + c * 32 + u * 8);
                }
            };
            if (EPI == kEpiResLN) {
                coop_load(set * (kChunks / 2), pre[0]);
                coop_load(set * (kChunks / 2) + 1, pre[1]);
            }
            const long long t0 = clock64();
            const bool ready = mbar_wait(&tmem_full_bar[acc], acc_phase);
            t_wait += clock64() - t0;
            if (!ready) {
                atomicExch(p.status, kStatusEpilogueTimeout);
                break;
            }
            tcgen05_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * Shape::kAccCols;
            const int m0 = m_blk * kBM;
            const int64_t m = (int64_t)m0 + row;
            uint32_t raw[32], h[32], l[32];
            float y[32];
            auto release_tmem = [&]() {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_cluster_relaxed(map_to_cta(&tmem_empty_bar[acc], 0));
                    else mbar_arrive(&tmem_empty_bar[acc]);
                }
            };

            if (EPI == kEpiF32) {
#pragma unroll 1
                for (int c = set; c < kChunks; c += 2) {
                    tmem_ld_32x32(t_acc + c * 32, raw);
                    tmem_wait_ld();
                    const int n0 = (p.group_cols ? n_blk * p.group_cols : n_blk * BN) + c * 32;
                    const int n_end = p.group_cols ? min(p.N, (n_blk + 1) * p.group_cols) : p.N;
                    if (n0 >= n_end) continue;   // warp-uniform
                    float* dst = p.out_f32 + m * p.ld_f32 + n0;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + j < n_end) {
                            const float v = __uint_as_float(raw[j]) * scale + __ldg(p.bias + n0 + j);
                            dst[j] = p.relu == 1 ? fmaxf(v, 0.f)
                                   : p.relu == 2 ? 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)) : v;
                        }
                }
                release_tmem();
            } else if (EPI == kEpiPlanes) {
                // set s owns columns [128 s, 128 s + 128) of the tile: two 64-column groups,
                // each leaves as one hi-plane and one lo-plane TMA store
#pragma unroll 1
                for (int g = 0; g < kChunks / 4; ++g) {
                    const int col = set * (BN / 2) + g * 64;
                    const int n0 = n_blk * BN + col;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        tmem_ld_32x32(t_acc + col + half * 32, raw);
                        if (p.debug_flags & 2) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) y[j] = 0.f;
                        } else {
                            load_params32(p.bias + n0 + half * 32, y);
                        }
                        if (p.split_acc) {   // + the correction accumulator (second TMEM buffer)
                            uint32_t corr[32];
                            tmem_ld_32x32(t_acc + Shape::kAccCols + col + half * 32, corr);
                            tmem_wait_ld();
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(corr[j]));
                        }
                        tmem_wait_ld();
                        if (g + 1 == kChunks / 4 && half == 1) release_tmem();   // last read
                        // the activation is launch-uniform: keep erff out of the ReLU / linear loop
                        if (p.relu == 2) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float v = fmaf(__uint_as_float(raw[j]), scale, y[j]);
                                y[j] = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
                            }
                        } else {
                            const float floor = p.relu == 1 ? 0.f : -INFINITY;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                y[j] = fmaxf(fmaf(__uint_as_float(raw[j]), scale, y[j]), floor);
                        }
                        split32(y, h, l, half * 16);
                    }
                    stage_store_plane(&map_out, stage, set, row, elected, h, n0, m0, 0, p.debug_flags);
                    // columns below `hi_only_cols` are consumed as single fp16 planes (attention Q / K):
                    // their lo plane is never read, so it is not written (warp-uniform)
                    if (n0 >= p.hi_only_cols)
                        stage_store_plane(&map_out, stage, set, row, elected, l, n0, m0, 1, p.debug_flags);
                }
            } else if (EPI == kEpiConvIn) {
                const SeqInfo s = p.seqs[p.tile_seq[m_blk]];
                const int t = (int)(m - s.row0);
                const bool in_tensor = t < s.tensor_len, valid = t < s.valid_len;
#pragma unroll 1
                for (int g = 0; g < kChunks / 4; ++g) {
                    const int col = set * (BN / 2) + g * 64;
                    const int n0 = n_blk * BN + col;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float pe[32];
                        tmem_ld_32x32(t_acc + col + half * 32, raw);
                        load_params32(p.bias + n0 + half * 32, y);
                        load_params32(p.pe + (int64_t)(in_tensor ? t : 0) * p.N + n0 + half * 32, pe);
                        tmem_wait_ld();
                        if (g + 1 == kChunks / 4 && half == 1) release_tmem();
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float v = valid ? fmaf(__uint_as_float(raw[j]), scale, y[j]) : 0.f;
                            y[j] = in_tensor ? v + pe[j] : 0.f;
                        }
                        split32(y, h, l, half * 16);
                    }
                    stage_store_plane(&map_out, stage, set, row, elected, h, n0, m0, 0, p.debug_flags);
                    stage_store_plane(&map_out, stage, set, row, elected, l, n0, m0, 1, p.debug_flags);
                }
            } else if (EPI == kEpiResLN) {
                // the tile is one full row block of the hidden state (BN == N == hidden);
                // a row is shared by the two threads (one per set) with the same `row`
                const SeqInfo s = p.seqs[p.tile_seq[m_blk]];
                const bool in_tensor = (int)(m - s.row0) < s.tensor_len;
                const long long ln_t0 = clock64();
                float sum = 0.f;
                // The residual tile comes from HBM / L2.  Row-per-thread loads would touch 32
                // different lines per instruction (16 bytes each); instead the 128 threads of
                // the set load each 32-column chunk cooperatively (8 rows x 64 contiguous
                // bytes per warp instruction), pass it through the set's staging tile
                // ([hi|lo][128 rows][64 B], XOR-swizzled) and read their own row back.  The
                // loads run two chunks ahead (the first two are issued before the accumulator wait).
                const int c_first = set * (kChunks / 2);   // set s owns chunks [4 s, 4 s + 4)
                const uint32_t stage_addr = smem_u32(stage);
#pragma unroll
                for (int i = 0; i < kChunks / 2; ++i) {
                    const int c = c_first + i;
                    if (i == 0 && elected) bulk_wait_read_all();   // previous tile's stores drained
                    named_bar_sync(1 + set, 128);                   // staging tile is free
#pragma unroll
                    for (int it8 = 0; it8 < 8; ++it8) {
                        const int idx = it8 * 128 + tid_in_set;
                        const int u = idx & 3, r = (idx >> 2) & 127, plane = idx >> 9;
                        st_shared_v4(stage_addr + (uint32_t)plane * 8192 + (uint32_t)r * 64 +
                                         (((uint32_t)u ^ ((uint32_t)(r >> 1) & 3)) << 4),
                                     pre[i & 1][it8].x, pre[i & 1][it8].y, pre[i & 1][it8].z, pre[i & 1][it8].w);
                    }
                    named_bar_sync(1 + set, 128);
                    if (i + 2 < kChunks / 2) coop_load(c + 2, pre[i & 1]);
                    tmem_ld_32x32(t_acc + c * 32, raw);
                    load_params32(p.bias + c * 32, y);
                    uint4 rh[4], rl[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t addr = stage_addr + (uint32_t)row * 64 +
                                              (((uint32_t)u ^ ((uint32_t)(row >> 1) & 3)) << 4);
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(rh[u].x), "=r"(rh[u].y), "=r"(rh[u].z), "=r"(rh[u].w)
                                     : "r"(addr));
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(rl[u].x), "=r"(rl[u].y), "=r"(rl[u].z), "=r"(rl[u].w)
                                     : "r"(addr + 8192));
                    }
                    tmem_wait_ld();
                    const uint32_t* rhw = reinterpret_cast<const uint32_t*>(rh);
                    const uint32_t* rlw = reinterpret_cast<const uint32_t*>(rl);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&rhw[j]));
                        const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&rlw[j]));
                        const float v0 = fmaf(__uint_as_float(raw[2 * j]), scale, y[2 * j]) + (fh.x + fl.x);
                        const float v1 = fmaf(__uint_as_float(raw[2 * j + 1]), scale, y[2 * j + 1]) + (fh.y + fl.y);
                        sum += v0 + v1;
                        raw[2 * j] = __float_as_uint(v0);
                        raw[2 * j + 1] = __float_as_uint(v1);
                    }
                    tmem_st_32x32(t_acc + c * 32, raw);
                }
                tmem_wait_st();
                const long long ln_t1 = clock64();
                ln_part[set][row] = sum;
                named_bar_sync(3, 256);
                const float mean = (ln_part[0][row] + ln_part[1][row]) * (1.f / BN);
                named_bar_sync(3, 256);   // both partners have read the sums
                const long long ln_t2 = clock64();
                float sq = 0.f;
#pragma unroll 1
                for (int c = c_first; c < c_first + kChunks / 2; ++c) {
                    tmem_ld_32x32(t_acc + c * 32, raw);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float d = __uint_as_float(raw[j]) - mean;
                        sq = fmaf(d, d, sq);
                    }
                }
                const long long ln_t3 = clock64();
                ln_part[set][row] = sq;
                named_bar_sync(3, 256);
                const float rstd = rsqrtf((ln_part[0][row] + ln_part[1][row]) * (1.f / BN) + p.eps);
                named_bar_sync(3, 256);   // ln_part is free for the next tile
                const long long ln_t4 = clock64();
#pragma unroll 1
                for (int g = 0; g < kChunks / 4; ++g) {
                    const int n0 = (c_first + 2 * g) * 32;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float gamma[32];
                        tmem_ld_32x32(t_acc + n0 + half * 32, raw);
                        load_params32(p.gamma + n0 + half * 32, gamma);
                        load_params32(p.beta + n0 + half * 32, y);
                        tmem_wait_ld();
                        if (g + 1 == kChunks / 4 && half == 1) release_tmem();
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float v = fmaf((__uint_as_float(raw[j]) - mean) * rstd, gamma[j], y[j]);
                            y[j] = in_tensor ? v : 0.f;
                        }
                        split32(y, h, l, half * 16);
                    }
                    stage_store_plane(&map_out, stage, set, row, elected, h, n0, m0, 0, p.debug_flags);
                    stage_store_plane(&map_out, stage, set, row, elected, l, n0, m0, 1, p.debug_flags);
                }
                if (p.trace_ln && warp == 2 && lane == 0) {
                    atomicAdd(p.trace_ln + 0, (unsigned long long)(ln_t1 - ln_t0));
                    atomicAdd(p.trace_ln + 1, (unsigned long long)(ln_t2 - ln_t1));
                    atomicAdd(p.trace_ln + 2, (unsigned long long)(ln_t3 - ln_t2));
                    atomicAdd(p.trace_ln + 3, (unsigned long long)(ln_t4 - ln_t3));
                    atomicAdd(p.trace_ln + 4, (unsigned long long)(clock64() - ln_t4));
                    atomicAdd(p.trace_ln + 5, 1ull);
                }
            } else if (EPI == kEpiConvOut) {
                // BN == 64 >= O: all channels of a frame live in one thread (set 0)
                if (set == 0) {
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
                    release_tmem();
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
                } else {
                    release_tmem();
                }
            }
        }
        if (elected) bulk_wait_all();   // smem must outlive the last TMA store
        if (p.trace && warp == 2 && lane == 0) {
            atomicAdd(p.trace + 3, (unsigned long long)t_wait);
            atomicAdd(p.trace + 4, (unsigned long long)(clock64() - t_begin));
        }
    }

    tcgen05_fence_before();
    if (PAIR) cluster_sync_all();   // the peer's smem / TMEM / barriers stay valid until both are done
    else __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        if (PAIR) tmem_dealloc_pair<Shape::kTmemCols>(tmem_base);
        else tmem_dealloc<Shape::kTmemCols>(tmem_base);
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
bool pdl_enabled() {
    static const bool on = [] { const char* v = getenv("PPGS_B200_PDL"); return v && atoi(v) != 0; }();
    return on;
}

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
                   uint32_t box_planes, uint32_t row_step) {
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
        box[0] = kBK; box[1] = box_rows * row_step; box[2] = box_planes;
        elem[1] = row_step;   // box_rows rows, every row_step-th one
        if (box[1] > 256) {
            set_error("make_plane_map: box of %u rows with step %u exceeds the TMA limit", box_rows, row_step);
            return PPGS_E_INVALID;
        }
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

int make_store_map(CUtensorMap* map, __half* base, uint64_t inner, uint64_t rows,
                   uint64_t plane_stride_elems) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return PPGS_E_CUDA;
    }
    cuuint64_t dims[3] = {inner, rows, 2};
    cuuint64_t strides[2] = {inner * 2, plane_stride_elems * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)kBM, 1}, elem[3] = {1, 1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, strides, box, elem,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (store map) failed with CUresult %d", (int)rc);
        return PPGS_E_CUDA;
    }
    return PPGS_OK;
}

template <int BN, int EPI, bool PAIR>
static int launch_one(ppgs_engine* e, const char* name, const CUtensorMap& map_a,
                      const CUtensorMap& map_b, const CUtensorMap& map_out, const GemmParams& p,
                      cudaStream_t stream) {
    using Shape = GemmShape<BN, PAIR>;
    static PerDeviceOnce attr;
    if (attr.first(e->device)) {
        PPGS_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, PAIR>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Shape::kSmemBytes));
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attrs[2];
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Shape::kSmemBytes;
    cfg.stream = stream;
    if (PAIR) {
        if (p.m_tiles % 2) {
            set_error("gemm_tc: the CTA-pair kernel needs an even number of row tiles");
            return PPGS_E_INVALID;
        }
        const int pair_tiles = (p.m_tiles / 2) * p.n_tiles;
        cfg.gridDim = dim3(2 * std::min(pair_tiles, e->sm_count / 2));
        attrs[0].id = cudaLaunchAttributeClusterDimension;
        attrs[0].val.clusterDim.x = 2;
        attrs[0].val.clusterDim.y = 1;
        attrs[0].val.clusterDim.z = 1;
        cfg.attrs = attrs;
        cfg.numAttrs = 1;
    } else {
        cfg.gridDim = dim3(std::min(p.m_tiles * p.n_tiles, e->sm_count));
    }
    if (pdl_enabled()) {
        cfg.attrs = attrs;
        attrs[cfg.numAttrs].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[cfg.numAttrs].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs += 1;
    }
    {
        LaunchScope scope(e, name, stream);
        PPGS_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EPI, PAIR>, map_a, map_b, map_out, p));
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

int launch_gemm_tc(ppgs_engine* e, const char* name, int bn, int epilogue,
                   const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap* map_out,
                   const GemmParams& p, cudaStream_t stream) {
    if (p.m_tiles <= 0 || p.n_tiles <= 0 || p.cblocks <= 0 || !p.scale || !p.status) {
        set_error("gemm_tc: bad parameters");
        return PPGS_E_INVALID;
    }
    const bool stores_planes = epilogue == kEpiPlanes || epilogue == kEpiConvIn || epilogue == kEpiResLN;
    if (stores_planes && !map_out) {
        set_error("gemm_tc: epilogue %d needs an output tensor map", epilogue);
        return PPGS_E_INVALID;
    }
    const CUtensorMap& out_map = map_out ? *map_out : map_a;   // unused when not storing planes
    const bool pair = p.pair != 0;
#define PPGS_GEMM_CASE(BN_, EPI_, PAIR_) \
    if (bn == BN_ && epilogue == EPI_ && pair == PAIR_) \
        return launch_one<BN_, EPI_, PAIR_>(e, name, map_a, map_b, out_map, p, stream)
    PPGS_GEMM_CASE(256, kEpiF32, false);
    PPGS_GEMM_CASE(256, kEpiF32, true);
    PPGS_GEMM_CASE(128, kEpiF32, false);
    PPGS_GEMM_CASE(64, kEpiF32, false);
    PPGS_GEMM_CASE(256, kEpiPlanes, false);
    PPGS_GEMM_CASE(256, kEpiPlanes, true);
    PPGS_GEMM_CASE(256, kEpiConvIn, false);
    PPGS_GEMM_CASE(256, kEpiConvIn, true);
    PPGS_GEMM_CASE(256, kEpiResLN, false);
    PPGS_GEMM_CASE(256, kEpiResLN, true);
    PPGS_GEMM_CASE(64, kEpiConvOut, false);
#undef PPGS_GEMM_CASE
    set_error("gemm_tc: no kernel for BN=%d epilogue=%d pair=%d", bn, epilogue, (int)pair);
    return PPGS_E_UNSUPPORTED;
}

}  // namespace tc
}  // namespace ppgs
