// Fused feed-forward block of torch.nn.TransformerEncoderLayer as used by
// ppgs/model/transformer.py:35-41 (post-LN, ReLU, dim_feedforward 2048):
//
//     x <- LayerNorm2(x + W2 relu(W1 x + b1) + b2)
//
// in ONE kernel per layer: the (rows x 2048) hidden activation never leaves the SM.  Unfused
// it is written and read back as split-fp16 planes, 8 KB per row and layer = 6.7 GB per
// 64 x 10 s batch — more than half of the step's HBM traffic, on a part that runs at its power
// cap (profiles/r02_power_probe.jsonl: the step without those stores is 28 % shorter).
//
// A CTA pair (cluster of 2, tcgen05 cta_group::2, M = 256) owns 256 rows; per 128-wide hidden
// chunk c (16 chunks):
//     G1_c   Hacc[c & 1] = X . W1_c^T        N = 128, K = 256   (TMEM cols 256 + 128 (c & 1))
//     E_c    relu(Hacc * s1 + b1) -> split fp16 -> H1 smem tile [kc 2][hi|lo][128 rows][64]
//     G2_c   Y += H1 . W2_c^T                N = 256, K = 128   (TMEM cols 0..255)
// issued as G1_0, G1_1, G2_0, G1_2, G2_1, ... so that the epilogue warps turn Hacc_c into H1
// while the tensor pipe runs G1_{c+1} / G2_{c-1}.  Round 1's version used 64-wide chunks: at
// N = 64 the MMAs re-read their 4 KB A slice every 32 cycles and shared memory, not the
// tensor pipe, set the pace (it was 40 % slower than the two GEMMs and off by default).  With
// N = 128 the operand reads need 96 B/clk of the 128 B/clk port.
//
// Everything streams through ONE ring of three 48 KB slots filled by one producer warp in the
// order the MMA warp consumes it: a G1 slot = X k-block (hi|lo, 32 KB) + this CTA's half of
// the W1 k-block (64 rows, hi|lo, 16 KB); a G2 slot = this CTA's half of the W2 k-block
// (128 rows, hi|lo, 32 KB).  X is re-read from L2 per chunk (the resident X tile of round 1
// would leave no room for a 128-wide H1 tile).
//
// LayerNorm epilogue (per tile): the Y accumulator is read ONCE into registers (128 values per
// thread, two threads per row) and released at once, so the next tile's MMAs start while
// bias + residual + LayerNorm + split + TMA stores run from registers.  The residual tile
// comes by TMA into the dead H1 tile.
//
// Warp roles (384 threads): 0 producer, 1 MMA issuer (leader CTA) + TMEM allocation, 2-3
// idle, 4-11 epilogue (two threads per row); setmaxnreg moves registers to the epilogue.
#include "gemm_tc.cuh"

namespace ppgs {
namespace tc {

namespace {

constexpr int kFfnThreads = 384;
constexpr int kFC = 128;                        // hidden-chunk width
constexpr int kTileBytes = 16384;               // [128 rows][64] fp16, 128-byte swizzle
constexpr int kSlotBytes = 3 * kTileBytes;      // 48 KB
constexpr int kStages = 3;
constexpr int kW1PlaneBytes = 64 * 128;         // [64 rows][64] = 8 KB
constexpr int kH1Off = kStages * kSlotBytes;    // [kc 2][hi|lo][128 rows][64] = 64 KB
constexpr size_t kFfnSmem = kH1Off + 4 * kTileBytes;   // 208 KB
constexpr int kYCol = 0, kHaccCol = 256;

__device__ __forceinline__ void load_row32(const float* src, float (&out)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + j);
        out[4 * j] = v.x;
        out[4 * j + 1] = v.y;
        out[4 * j + 2] = v.z;
        out[4 * j + 3] = v.w;
    }
}

// one 128-byte row of a 128B-swizzled [128 rows][64 fp16] tile <- 32 packed words
__device__ __forceinline__ void stage_row128(uint32_t tile, int row, const uint32_t (&w)[32]) {
    const uint32_t base = tile + (uint32_t)row * 128, sw = (uint32_t)row & 7;
#pragma unroll
    for (int u = 0; u < 8; ++u)
        st_shared_v4(base + (((uint32_t)u ^ sw) << 4), w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
}

__device__ __forceinline__ void load_row128(uint32_t tile, int row, uint32_t (&w)[32]) {
    const uint32_t base = tile + (uint32_t)row * 128, sw = (uint32_t)row & 7;
#pragma unroll
    for (int u = 0; u < 8; ++u)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(w[4 * u]), "=r"(w[4 * u + 1]), "=r"(w[4 * u + 2]), "=r"(w[4 * u + 3])
                     : "r"(base + (((uint32_t)u ^ sw) << 4)));
}


// pair-tile index -> row tile of this CTA (identity unless a row-tile window is set)
__device__ __forceinline__ int pair_row_tile(const FfnParams& p, int pt, int rank) {
    if (p.reverse) pt = p.m_tiles / 2 - 1 - pt;
    if (p.win_size == 0) return 2 * pt + rank;
    const int per_seq = p.win_size / 2;
    const int seq = pt / per_seq, u = pt - seq * per_seq;
    return seq * p.win_stride + p.seqs[seq].src_start + 2 * u + rank;
}

// LayerNorm epilogue shared by the fused FFN and the projection kernel.  Thread = (row, set):
// 128 accumulator columns [128 set, 128 set + 128) of one of the CTA's 128 rows.
//   1. accumulator -> registers, then `release_bar` (cluster address) is signalled: the MMA
//      warp may overwrite the TMEM columns while everything below runs from registers
//   2. + bias + residual; the residual rows arrive by TMA in the set's 32 KB staging region
//      (one 64-column group, hi|lo, at a time; `res_full` phases continue from `res_phase`)
//   3. two-pass mean / variance over the 256 columns (partner thread through `ln_part`)
//   4. normalise, zero rows outside the sequence tensor, split, stage, TMA store
// Returns false on a barrier time-out.  All 256 epilogue threads must call it together.
__device__ __forceinline__ bool residual_layernorm_store(
    uint32_t t_acc, uint32_t release_bar, float scale, const float* bias, const float* gamma_p,
    const float* beta_p, float eps, const SeqInfo* seqs, const int* tile_seq, int m_blk, int row, int set,
    int lane, bool elected, unsigned char* stage_smem, uint64_t* res_full, uint32_t res_phase,
    float (*ln_part)[kBM], const CUtensorMap* map_res, const CUtensorMap* map_out,
    bool res0_issued = false, uint64_t res_policy = kL2EvictNormal) {
    const int m0 = m_blk * kBM;
    const int64_t m = (int64_t)m0 + row;
    bool ok = true;
    float v[128];
    {
        uint32_t raw[32];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            tmem_ld_32x32(t_acc + i * 32, raw);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[i * 32 + j] = __uint_as_float(raw[j]) * scale;
        }
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster_relaxed(release_bar);   // the next tile may overwrite the accumulator
    // residual rows: TMA into the dead H1 tile, one 64-column group (hi|lo, 32 KB) per set
    // at a time; y_full implies the last G2 has read H1
    const uint32_t res_tile = smem_u32(stage_smem) + set * 2 * kTileBytes;
    float sum = 0.f;
#pragma unroll
    for (int gq = 0; gq < 2; ++gq) {
        // (the caller may have requested group 0 already, under the wait for the accumulator)
        if (gq > 0 || !res0_issued) {
            named_bar_sync(1 + set, 128);   // previous group consumed by all rows of the set
            if (elected) {
                mbar_arrive_expect_tx(res_full, 2 * kTileBytes);
                tma_load_3d_hint(stage_smem + set * 2 * kTileBytes, map_res, res_full, set * 128 + gq * 64, m0, 0,
                                 res_policy);
            }
        }
        float b[32];
        uint32_t rh[32], rl[32];
        if (!mbar_wait(res_full, (res_phase + gq) & 1)) { ok = false; break; }
        load_row128(res_tile, row, rh);
        load_row128(res_tile + kTileBytes, row, rl);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            load_row32(bias + set * 128 + gq * 64 + half * 32, b);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&rh[half * 16 + j]));
                const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&rl[half * 16 + j]));
                const int at = gq * 64 + half * 32 + 2 * j;
                v[at] = (v[at] + b[2 * j]) + (fh.x + fl.x);
                v[at + 1] = (v[at + 1] + b[2 * j + 1]) + (fh.y + fl.y);
                sum += v[at] + v[at + 1];
            }
        }
    }
    if (!ok) return false;
    ln_part[set][row] = sum;
    named_bar_sync(3, 256);
    const float mean = (ln_part[0][row] + ln_part[1][row]) * (1.f / 256);
    named_bar_sync(3, 256);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 128; ++i) {
        const float d = v[i] - mean;
        sq = fmaf(d, d, sq);
    }
    ln_part[set][row] = sq;
    named_bar_sync(3, 256);
    const float rstd = rsqrtf((ln_part[0][row] + ln_part[1][row]) * (1.f / 256) + eps);
    named_bar_sync(3, 256);   // also: every row of both sets has consumed its residual tile
    const SeqInfo s = seqs[tile_seq[m_blk]];
    const bool in_tensor = (int)(m - s.row0) < s.tensor_len;
    // normalise + split + stage (hi -> first 16 KB tile of the set, lo -> second) + TMA store
#pragma unroll
    for (int gq = 0; gq < 2; ++gq) {
        uint32_t h[32], l[32];
        float gamma[32], beta[32];
        const int n0 = set * 128 + gq * 64;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            load_row32(gamma_p + n0 + half * 32, gamma);
            load_row32(beta_p + n0 + half * 32, beta);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int at = gq * 64 + half * 32 + 2 * j;
                const float v0 = fmaf((v[at] - mean) * rstd, gamma[2 * j], beta[2 * j]);
                const float v1 = fmaf((v[at + 1] - mean) * rstd, gamma[2 * j + 1], beta[2 * j + 1]);
                split2_f16(in_tensor ? v0 : 0.f, in_tensor ? v1 : 0.f, h[half * 16 + j], l[half * 16 + j]);
            }
        }
        if (gq == 1) {
            if (elected) bulk_wait_read_all();   // group 0's stores have drained the tiles
            named_bar_sync(1 + set, 128);
        }
        stage_row128(res_tile, row, h);
        stage_row128(res_tile + kTileBytes, row, l);
        fence_proxy_async_smem();
        named_bar_sync(1 + set, 128);
        if (elected) {
            tma_store_3d(map_out, stage_smem + set * 2 * kTileBytes, n0, m0, 0);
            tma_store_3d(map_out, stage_smem + set * 2 * kTileBytes + kTileBytes, n0, m0, 1);
            bulk_commit_group();
        }
    }
    return ok;
}

__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_out,
                 const __grid_constant__ CUtensorMap map_res, const FfnParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* h1_smem = smem + kH1Off;
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages];
    __shared__ __align__(8) uint64_t hacc_full[2], hacc_empty[2], h1_full, h1_empty, y_full, y_empty;
    __shared__ __align__(8) uint64_t res_full[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float ln_part[2][kBM];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int worker = (int)(blockIdx.x >> 1), workers = (int)(gridDim.x >> 1);
    const int pair_tiles = p.m_tiles / 2, NC = p.num_chunks;
    if (smem_u32(smem) & 1023u) {
        if (threadIdx.x == 0) atomicExch(p.status, kStatusBadAlignment);
        return;
    }

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&hacc_full[i], 1);
            mbar_init(&hacc_empty[i], 16);   // epilogue warps of both CTAs
            mbar_init(&res_full[i], 1);
        }
        mbar_init(&h1_full, 16);
        mbar_init(&h1_empty, 1);
        mbar_init(&y_full, 1);
        mbar_init(&y_empty, 16);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair<512>(&tmem_slot);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (warp == 0) {
            // ------------------------------------------------ producer: one ring, MMA order
            if (lane == 0) {
                prefetch_tensormap(&map_x);
                prefetch_tensormap(&map_w1);
                prefetch_tensormap(&map_w2);
                const uint32_t g1_tx = 2u * (p.planes * kTileBytes + p.planes * kW1PlaneBytes);
                const uint32_t g2_tx = 2u * p.planes * kTileBytes;
                int stage = 0;
                uint32_t phase = 0;
                bool ok = true;
                auto acquire = [&](uint32_t tx) -> unsigned char* {
                    if (!mbar_wait(&empty_bar[stage], phase ^ 1)) {
                        ok = false;
                        return nullptr;
                    }
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx);
                    return smem + stage * kSlotBytes;
                };
                auto advance = [&]() {
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                };
                for (int pt = worker; pt < pair_tiles && ok; pt += workers) {
                    const int m_blk = pair_row_tile(p, pt, (int)rank);
                    auto load_g1 = [&](int c) {
                        for (int kb = 0; kb < 4 && ok; ++kb) {
                            unsigned char* slot = acquire(g1_tx);
                            if (!slot) return;
                            const uint32_t bar = map_to_cta(&full_bar[stage], 0);
                            tma_load_3d_pair(slot, &map_x, bar, kb * kBK, m_blk * kBM, 0);
                            tma_load_4d_pair(slot + 2 * kTileBytes, &map_w1, bar, kb * kBK, c * kFC + (int)rank * 64, 0, 0);
                            advance();
                        }
                    };
                    auto load_g2 = [&](int c) {
                        for (int kb = 0; kb < 2 && ok; ++kb) {
                            unsigned char* slot = acquire(g2_tx);
                            if (!slot) return;
                            const uint32_t bar = map_to_cta(&full_bar[stage], 0);
                            tma_load_4d_pair(slot, &map_w2, bar, c * kFC + kb * kBK, (int)rank * 128, 0, 0);
                            advance();
                        }
                    };
                    load_g1(0);
                    for (int c = 0; c < NC && ok; ++c) {
                        if (c + 1 < NC) load_g1(c + 1);
                        load_g2(c);
                    }
                }
                if (!ok) atomicExch(p.status, kStatusProducerTimeout);
            }
        } else if (warp == 1) {
            // ------------------------------------------------ MMA issuer (leader CTA)
            if (lane == 0 && rank == 0) {
                constexpr uint32_t idesc1 = make_idesc_f16(2 * kBM, kFC);
                constexpr uint32_t idesc2 = make_idesc_f16(2 * kBM, 256);
                const uint32_t h1_addr = smem_u32(h1_smem);
                const bool two = p.planes == 2;
                int stage = 0;
                uint32_t phase = 0;
                bool ok = true;
                uint32_t g = 0;    // chunks issued by G1 so far (all tiles)
                uint32_t g2 = 0;   // chunks issued by G2 so far
                int it = 0;
                auto next_slot = [&]() -> uint32_t {
                    if (!mbar_wait(&full_bar[stage], phase)) {
                        ok = false;
                        return 0;
                    }
                    tcgen05_fence_after();
                    return smem_u32(smem + stage * kSlotBytes);
                };
                auto release_slot = [&]() {
                    umma_commit_pair(&empty_bar[stage]);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                };
                auto issue_g1 = [&]() {
                    const int hb = (int)(g & 1);
                    if (!mbar_wait(&hacc_empty[hb], ((g >> 1) & 1) ^ 1)) { ok = false; return; }
                    tcgen05_fence_after();
                    const uint32_t d = tmem_base + kHaccCol + hb * kFC;
                    for (int kb = 0; kb < 4 && ok; ++kb) {
                        const uint32_t a0 = next_slot();
                        if (!ok) return;
                        const uint32_t b0 = a0 + 2 * kTileBytes;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t koff = k * 32;
                            const uint64_t da0 = smem_desc_kmajor_sw128(a0 + koff);
                            const uint64_t db0 = smem_desc_kmajor_sw128(b0 + koff);
                            umma_f16_pair(d, da0, db0, idesc1, (kb > 0 || k > 0) ? 1u : 0u);
                            if (two) {
                                umma_f16_pair(d, da0, smem_desc_kmajor_sw128(b0 + kW1PlaneBytes + koff), idesc1, 1);
                                umma_f16_pair(d, smem_desc_kmajor_sw128(a0 + kTileBytes + koff), db0, idesc1, 1);
                            }
                        }
                        release_slot();
                    }
                    umma_commit_pair(&hacc_full[hb]);
                    ++g;
                };
                auto issue_g2 = [&](int c) {
                    if (c == 0 && !mbar_wait(&y_empty, (uint32_t)(it & 1) ^ 1)) { ok = false; return; }
                    if (!mbar_wait(&h1_full, g2 & 1)) { ok = false; return; }
                    tcgen05_fence_after();
                    const uint32_t d = tmem_base + kYCol;
                    for (int kb = 0; kb < 2 && ok; ++kb) {
                        const uint32_t b0 = next_slot();
                        if (!ok) return;
                        const uint32_t a0 = h1_addr + kb * 2 * kTileBytes;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t koff = k * 32;
                            const uint64_t da0 = smem_desc_kmajor_sw128(a0 + koff);
                            const uint64_t db0 = smem_desc_kmajor_sw128(b0 + koff);
                            umma_f16_pair(d, da0, db0, idesc2, (c > 0 || kb > 0 || k > 0) ? 1u : 0u);
                            if (two) {
                                umma_f16_pair(d, da0, smem_desc_kmajor_sw128(b0 + kTileBytes + koff), idesc2, 1);
                                umma_f16_pair(d, smem_desc_kmajor_sw128(a0 + kTileBytes + koff), db0, idesc2, 1);
                            }
                        }
                        release_slot();
                    }
                    umma_commit_pair(&h1_empty);
                    ++g2;
                };
                for (int pt = worker; pt < pair_tiles && ok; pt += workers, ++it) {
                    issue_g1();
                    for (int c = 0; c < NC && ok; ++c) {
                        if (c + 1 < NC) issue_g1();
                        if (ok) issue_g2(c);
                    }
                    if (ok) umma_commit_pair(&y_full);
                }
                if (!ok) atomicExch(p.status, kStatusMmaTimeout);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        // ------------------------------------------------ epilogue warps 4..11
        const int set = (warp - 4) >> 2;          // 64-column half of a hidden chunk / 128-column half of Y
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const bool elected = quad == 0 && lane == 0;          // one thread per set
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t h1_addr = smem_u32(h1_smem);
        const uint32_t hacc_empty_remote[2] = {map_to_cta(&hacc_empty[0], 0), map_to_cta(&hacc_empty[1], 0)};
        const uint32_t h1_full_remote = map_to_cta(&h1_full, 0);
        const uint32_t y_empty_remote = map_to_cta(&y_empty, 0);
        const float scale1 = *p.scale1, scale2 = *p.scale2;
        bool ok = true;
        int it = 0;
        uint32_t g = 0;
        for (int pt = worker; pt < pair_tiles && ok; pt += workers, ++it) {
            const int m_blk = pair_row_tile(p, pt, (int)rank);
            const int m0 = m_blk * kBM;
            const int64_t m = (int64_t)m0 + row;
            // ---- hidden chunks: Hacc -> relu -> split -> H1 (A operand of G2); this thread owns
            // 64 columns (kc = set) of its row
#pragma unroll 1
            for (int c = 0; c < NC; ++c, ++g) {
                const int hb = (int)(g & 1);
                uint32_t raw[2][32], h[32], l[32];
                float b[32];
                if (!mbar_wait(&hacc_full[hb], (g >> 1) & 1)) { ok = false; break; }
                tcgen05_fence_after();
                const uint32_t col = kHaccCol + hb * kFC + set * 64;
                tmem_ld_32x32(t_lane + col, raw[0]);
                tmem_ld_32x32(t_lane + col + 32, raw[1]);
                load_row32(p.bias1 + c * kFC + set * 64, b);
                tmem_wait_ld();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(hacc_empty_remote[hb]);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    split2_f16(fmaxf(fmaf(__uint_as_float(raw[0][2 * j]), scale1, b[2 * j]), 0.f),
                               fmaxf(fmaf(__uint_as_float(raw[0][2 * j + 1]), scale1, b[2 * j + 1]), 0.f), h[j], l[j]);
                load_row32(p.bias1 + c * kFC + set * 64 + 32, b);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    split2_f16(fmaxf(fmaf(__uint_as_float(raw[1][2 * j]), scale1, b[2 * j]), 0.f),
                               fmaxf(fmaf(__uint_as_float(raw[1][2 * j + 1]), scale1, b[2 * j + 1]), 0.f), h[16 + j],
                               l[16 + j]);
                // the H1 tile is free once G2 of the previous chunk has retired
                if (!mbar_wait(&h1_empty, (g & 1) ^ 1)) { ok = false; break; }
                stage_row128(h1_addr + set * 2 * kTileBytes, row, h);
                stage_row128(h1_addr + set * 2 * kTileBytes + kTileBytes, row, l);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(h1_full_remote);
            }
            if (!ok) break;
            // ---- LayerNorm epilogue: Y -> registers (and released), + bias2 + residual
            if (!mbar_wait(&y_full, (uint32_t)(it & 1))) { ok = false; break; }
            tcgen05_fence_after();
            if (!residual_layernorm_store(t_lane + kYCol + set * 128, y_empty_remote, scale2, p.bias2, p.gamma, p.beta,
                                          p.eps, p.seqs, p.tile_seq, m_blk, row, set, lane, elected, h1_smem,
                                          &res_full[set], (uint32_t)(2 * it), ln_part, &map_res, &map_out, false,
                                          p.dead_policy)) {
                ok = false;
                break;
            }
            // the staging tiles alias H1: drain the stores before the next tile's chunks
            if (elected) bulk_wait_read_all();
            named_bar_sync(3, 256);
        }
        if (elected) bulk_wait_all();
        if (!ok) atomicExch(p.status, kStatusEpilogueTimeout);
    }

    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_pair<512>(tmem_base);
    }
}


// ---------------------------------------------------------------------------
// x <- LayerNorm(x + A . W^T + b) for a hidden-256 projection (the attention out-projection,
// K = 256; any K % 64 == 0): the CTA-pair GEMM of gemm_tc.cu with the register-resident
// LayerNorm epilogue above.  The gemm_tc ResLN epilogue made three TMEM passes with the
// residual fetched cooperatively through one 16 KB staging tile and took 22 k cycles per tile
// against 6 k cycles of MMA at K = 256; here the accumulator (double-buffered, 2 x 256 TMEM
// columns) is released after one read, so the tensor pipe runs tile i+1 under the epilogue of
// tile i.  Ring: two 64 KB slots (A k-block hi|lo + this CTA's half of the W k-block hi|lo).
// ---------------------------------------------------------------------------
constexpr int kProjSlotBytes = 4 * kTileBytes;          // 64 KB
constexpr int kProjStages = 2;
constexpr int kProjStageOff = kProjStages * kProjSlotBytes;
constexpr size_t kProjSmem = kProjStageOff + 4 * kTileBytes;   // 192 KB

__global__ void __launch_bounds__(kFfnThreads, 1)
proj_ln_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
               const FfnParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* stage_smem = smem + kProjStageOff;
    __shared__ __align__(8) uint64_t full_bar[kProjStages], empty_bar[kProjStages];
    __shared__ __align__(8) uint64_t y_full[2], y_empty[2], res_full[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float ln_part[2][kBM];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int worker = (int)(blockIdx.x >> 1), workers = (int)(gridDim.x >> 1);
    const int pair_tiles = p.m_tiles / 2, KB = p.num_chunks;   // k-blocks of 64
    if (smem_u32(smem) & 1023u) {
        if (threadIdx.x == 0) atomicExch(p.status, kStatusBadAlignment);
        return;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < kProjStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&y_full[i], 1);
            mbar_init(&y_empty[i], 16);
            mbar_init(&res_full[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair<512>(&tmem_slot);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (warp == 0) {
            if (lane == 0) {
                prefetch_tensormap(&map_a);
                prefetch_tensormap(&map_w);
                const uint32_t tx = 2u * 2u * p.planes * kTileBytes;
                int stage = 0;
                uint32_t phase = 0;
                bool ok = true;
                for (int pt = worker; pt < pair_tiles && ok; pt += workers) {
                    const int m_blk = pair_row_tile(p, pt, (int)rank);
                    for (int kb = 0; kb < KB; ++kb) {
                        if (!mbar_wait(&empty_bar[stage], phase ^ 1)) { ok = false; break; }
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx);
                        unsigned char* slot = smem + stage * kProjSlotBytes;
                        const uint32_t bar = map_to_cta(&full_bar[stage], 0);
                        tma_load_3d_pair_hint(slot, &map_a, bar, kb * kBK, m_blk * kBM, 0, p.dead_policy);
                        tma_load_4d_pair(slot + 2 * kTileBytes, &map_w, bar, kb * kBK, (int)rank * 128, 0, 0);
                        if (++stage == kProjStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
                if (!ok) atomicExch(p.status, kStatusProducerTimeout);
            }
        } else if (warp == 1) {
            if (lane == 0 && rank == 0) {
                constexpr uint32_t idesc = make_idesc_f16(2 * kBM, 256);
                const bool two = p.planes == 2;
                int stage = 0;
                uint32_t phase = 0;
                bool ok = true;
                int it = 0;
                for (int pt = worker; pt < pair_tiles && ok; pt += workers, ++it) {
                    const int acc = it & 1;
                    if (!mbar_wait(&y_empty[acc], (uint32_t)((it >> 1) & 1) ^ 1)) { ok = false; break; }
                    tcgen05_fence_after();
                    const uint32_t d = tmem_base + acc * 256;
                    for (int kb = 0; kb < KB; ++kb) {
                        if (!mbar_wait(&full_bar[stage], phase)) { ok = false; break; }
                        tcgen05_fence_after();
                        const uint32_t a0 = smem_u32(smem + stage * kProjSlotBytes), b0 = a0 + 2 * kTileBytes;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t koff = k * 32;
                            const uint64_t da0 = smem_desc_kmajor_sw128(a0 + koff);
                            const uint64_t db0 = smem_desc_kmajor_sw128(b0 + koff);
                            umma_f16_pair(d, da0, db0, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                            if (two) {
                                umma_f16_pair(d, da0, smem_desc_kmajor_sw128(b0 + kTileBytes + koff), idesc, 1);
                                umma_f16_pair(d, smem_desc_kmajor_sw128(a0 + kTileBytes + koff), db0, idesc, 1);
                            }
                        }
                        umma_commit_pair(&empty_bar[stage]);
                        if (++stage == kProjStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    if (ok) umma_commit_pair(&y_full[acc]);
                }
                if (!ok) atomicExch(p.status, kStatusMmaTimeout);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        const int set = (warp - 4) >> 2, quad = warp & 3, row = quad * 32 + lane;
        const bool elected = quad == 0 && lane == 0;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t y_empty_remote[2] = {map_to_cta(&y_empty[0], 0), map_to_cta(&y_empty[1], 0)};
        const float scale = *p.scale2;
        bool ok = true;
        int it = 0;
        for (int pt = worker; pt < pair_tiles && ok; pt += workers, ++it) {
            const int m_blk = pair_row_tile(p, pt, (int)rank), acc = it & 1;
            // the first residual group of this tile is requested BEFORE the wait for the accumulator
            // (it depends on x only): its latency hides behind the tile's MMAs.  The staging tiles are
            // free once the previous tile's stores have read them and every row of the set is past them.
            named_bar_sync(1 + set, 128);
            if (elected) {
                bulk_wait_read_all();
                mbar_arrive_expect_tx(&res_full[set], 2 * kTileBytes);
                tma_load_3d_hint(stage_smem + set * 2 * kTileBytes, &map_res, &res_full[set], set * 128, m_blk * kBM, 0,
                                 p.dead_policy);
            }
            if (!mbar_wait(&y_full[acc], (uint32_t)((it >> 1) & 1))) { ok = false; break; }
            tcgen05_fence_after();
            if (!residual_layernorm_store(t_lane + acc * 256 + set * 128, y_empty_remote[acc], scale, p.bias2, p.gamma,
                                          p.beta, p.eps, p.seqs, p.tile_seq, m_blk, row, set, lane, elected,
                                          stage_smem, &res_full[set], (uint32_t)(2 * it), ln_part, &map_res,
                                          &map_out, true, p.dead_policy))
                ok = false;
        }
        if (elected) bulk_wait_all();
        if (!ok) atomicExch(p.status, kStatusEpilogueTimeout);
    }

    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_pair<512>(tmem_base);
    }
}

}  // namespace

int launch_ffn_fused(ppgs_engine* e, const CUtensorMap& map_x, const CUtensorMap& map_w1,
                     const CUtensorMap& map_w2, const CUtensorMap& map_out, const CUtensorMap& map_res,
                     const FfnParams& p, cudaStream_t stream) {
    if (p.m_tiles <= 0 || p.m_tiles % 2 || p.num_chunks <= 0) {
        set_error("ffn_fused: needs an even number of row tiles");
        return PPGS_E_INVALID;
    }
    static PerDeviceOnce attr;
    if (attr.first(e->device)) {
        PPGS_CUDA(cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kFfnSmem));
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attrs[2];
    cfg.blockDim = dim3(kFfnThreads);
    cfg.dynamicSmemBytes = kFfnSmem;
    cfg.stream = stream;
    cfg.gridDim = dim3(2 * std::min(p.m_tiles / 2, e->sm_count / 2));
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = 2;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    if (pdl_enabled()) {
        attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
    }
    {
        LaunchScope scope(e, "tc_ffn_fused_ln", stream);
        PPGS_CUDA(cudaLaunchKernelEx(&cfg, ffn_fused_kernel, map_x, map_w1, map_w2, map_out, map_res, p));
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}


int launch_proj_ln(ppgs_engine* e, const char* name, const CUtensorMap& map_a, const CUtensorMap& map_w,
                   const CUtensorMap& map_out, const CUtensorMap& map_res, const FfnParams& p,
                   cudaStream_t stream) {
    if (p.m_tiles <= 0 || p.m_tiles % 2 || p.num_chunks <= 0) {
        set_error("proj_ln: needs an even number of row tiles");
        return PPGS_E_INVALID;
    }
    static PerDeviceOnce attr;
    if (attr.first(e->device)) {
        PPGS_CUDA(cudaFuncSetAttribute(proj_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kProjSmem));
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attrs[2];
    cfg.blockDim = dim3(kFfnThreads);
    cfg.dynamicSmemBytes = kProjSmem;
    cfg.stream = stream;
    cfg.gridDim = dim3(2 * std::min(p.m_tiles / 2, e->sm_count / 2));
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = 2;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    if (pdl_enabled()) {
        attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
    }
    {
        LaunchScope scope(e, name, stream);
        PPGS_CUDA(cudaLaunchKernelEx(&cfg, proj_ln_kernel, map_a, map_w, map_out, map_res, p));
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

}  // namespace tc
}  // namespace ppgs
