// Fused feed-forward block of torch.nn.TransformerEncoderLayer as used by
// ppgs/model/transformer.py:35-41 (post-LN, ReLU, dim_feedforward 2048):
//
//     x <- LayerNorm2(x + W2 relu(W1 x + b1) + b2)
//
// in ONE kernel per layer: the (rows x 2048) hidden activation never leaves the SM.
// Unfused it costs 0.6 GB of HBM writes and 0.75 GB of reads per layer at the bench
// shape, and those stores were what bounded the linear1 GEMM.
//
// A CTA pair (cluster of 2, tcgen05 cta_group::2, M = 256) owns 256 rows:
//   X tile       [4 k-blocks][hi|lo][128 rows][64]  128 KB per CTA, loaded once per tile
//   per 64-wide hidden chunk c (32 chunks):
//     G1_c   Hacc[c&1] = X . W1_c^T          N = 64,  K = 256   (TMEM cols 256 + 64 (c&1))
//     E_c    relu(Hacc * s1 + b1) -> split fp16 -> H1 smem tile [hi|lo][128][64] (A operand)
//     G2_c   Y += H1 . W2_c^T                N = 256, K = 64    (TMEM cols 0..255)
//   LayerNorm epilogue on Y (+ b2 + residual), planes out through TMA stores
// The MMA issuer runs G1_{c+1} before G2_c so the tensor pipe is busy while the
// epilogue warps turn Hacc_c into H1.  Weights stream through small rings: four 8 KB
// W1 k-block slots and one 32 KB W2 slot per CTA (each CTA holds half of the N rows).
//
// Warp roles (352 threads): 0 X + W1 producer, 1 MMA issuer (leader CTA) + TMEM
// allocation, 2-9 epilogue (two threads per row), 10 W2 producer.
#include "gemm_tc.cuh"

namespace ppgs {
namespace tc {

constexpr int kFfnThreads = 352;
constexpr int kFC = 64;                        // hidden-chunk width
constexpr int kTileBytes = 16384;              // [128 rows][64] fp16
constexpr int kXBytes = 4 * 2 * kTileBytes;    // 128 KB
constexpr int kW1SubBytes = 2 * 32 * 128;      // [hi|lo][32 rows][64] = 8 KB
constexpr int kW1Bytes = 4 * kW1SubBytes;      // 32 KB
constexpr int kW2Bytes = 2 * kTileBytes;       // [hi|lo][128 rows][64] = 32 KB
constexpr int kH1Bytes = 2 * kTileBytes;       // 32 KB
constexpr size_t kFfnSmem = kXBytes + kW1Bytes + kW2Bytes + kH1Bytes;   // 224 KB
constexpr int kYCol = 0, kHaccCol = 256;

__device__ __forceinline__ bool timed_wait(uint64_t* bar, uint32_t parity, long long& acc) {
    const long long t0 = clock64();
    const bool ok = mbar_wait(bar, parity);
    acc += clock64() - t0;
    return ok;
}

__device__ __forceinline__ void stage_row128(uint32_t tile, int row, const uint32_t (&w)[32]) {
    const uint32_t base = tile + (uint32_t)row * 128, sw = (uint32_t)row & 7;
#pragma unroll
    for (int u = 0; u < 8; ++u)
        st_shared_v4(base + (((uint32_t)u ^ sw) << 4), w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
}

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

__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_out,
                 const FfnParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* x_smem = smem;
    unsigned char* w1_smem = x_smem + kXBytes;
    unsigned char* w2_smem = w1_smem + kW1Bytes;
    unsigned char* h1_smem = w2_smem + kW2Bytes;
    __shared__ __align__(8) uint64_t x_full, x_empty, w2_full, w2_empty, h1_full, h1_empty, y_full, y_empty;
    __shared__ __align__(8) uint64_t w1_full[4], w1_empty[4], hacc_full[2], hacc_empty[2];
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
        mbar_init(&x_full, 1);
        mbar_init(&x_empty, 8);     // this CTA's epilogue warps, after the residual was read from X
        mbar_init(&w2_full, 1);
        mbar_init(&w2_empty, 1);
        mbar_init(&h1_full, 2);     // one elected epilogue thread per CTA
        mbar_init(&h1_empty, 1);
        mbar_init(&y_full, 1);
        mbar_init(&y_empty, 16);
        for (int i = 0; i < 4; ++i) {
            mbar_init(&w1_full[i], 1);
            mbar_init(&w1_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&hacc_full[i], 1);
            mbar_init(&hacc_empty[i], 8);   // the 4 warps of the owning epilogue set, both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair<512>(&tmem_slot);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ producer: X tile + W1 chunks
        if (lane == 0) {
            prefetch_tensormap(&map_x);
            prefetch_tensormap(&map_w1);
            bool ok = true;
            int it = 0;
            for (int pt = worker; pt < pair_tiles && ok; pt += workers, ++it) {
                const int m_blk = 2 * pt + (int)rank;
                if (!mbar_wait(&x_empty, (uint32_t)(it & 1) ^ 1)) { ok = false; break; }
                if (rank == 0) mbar_arrive_expect_tx(&x_full, 2u * p.planes * 4 * kTileBytes);
                const uint32_t x_bar = map_to_cta(&x_full, 0);
                for (int kb = 0; kb < 4; ++kb)
                    tma_load_3d_pair(x_smem + kb * 2 * kTileBytes, &map_x, x_bar, kb * kBK, m_blk * kBM, 0);
                for (int c = 0; c < NC && ok; ++c) {
                    const uint32_t g = (uint32_t)(it * NC + c);
                    for (int kb = 0; kb < 4; ++kb) {
                        if (!mbar_wait(&w1_empty[kb], (g & 1) ^ 1)) { ok = false; break; }
                        if (rank == 0) mbar_arrive_expect_tx(&w1_full[kb], 2u * p.planes * 32 * 128);
                        tma_load_4d_pair(w1_smem + kb * kW1SubBytes, &map_w1, map_to_cta(&w1_full[kb], 0),
                                         kb * kBK, c * kFC + (int)rank * 32, 0, 0);
                    }
                }
            }
            if (!ok) atomicExch(p.status, kStatusProducerTimeout);
        }
    } else if (warp == 10) {
        // ------------------------------------------------ producer: W2 chunks
        if (lane == 0) {
            prefetch_tensormap(&map_w2);
            bool ok = true;
            int it = 0;
            for (int pt = worker; pt < pair_tiles && ok; pt += workers, ++it) {
                for (int c = 0; c < NC; ++c) {
                    const uint32_t g = (uint32_t)(it * NC + c);
                    if (!mbar_wait(&w2_empty, (g & 1) ^ 1)) { ok = false; break; }
                    if (rank == 0) mbar_arrive_expect_tx(&w2_full, 2u * p.planes * kTileBytes);
                    tma_load_4d_pair(w2_smem, &map_w2, map_to_cta(&w2_full, 0), c * kFC,
                                     (int)rank * 128, 0, 0);
                }
            }
            if (!ok) atomicExch(p.status, kStatusProducerTimeout);
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (leader CTA)
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc1 = make_idesc_f16(2 * kBM, kFC);
            constexpr uint32_t idesc2 = make_idesc_f16(2 * kBM, 256);
            const uint32_t x_addr = smem_u32(x_smem), w1_addr = smem_u32(w1_smem);
            const uint32_t w2_addr = smem_u32(w2_smem), h1_addr = smem_u32(h1_smem);
            const bool two = p.planes == 2;
            bool ok = true;
            int it = 0;
            long long tw[6] = {0, 0, 0, 0, 0, 0};
            const long long t_begin = clock64();
            // G2 of chunk (gg = global index, cc = index inside the tile)
            auto issue_g2 = [&](uint32_t gg, int cc) -> bool {
                if (cc == 0 && !timed_wait(&y_empty, (uint32_t)(it & 1) ^ 1, tw[5])) return false;
                if (!timed_wait(&h1_full, gg & 1, tw[3])) return false;
                if (!timed_wait(&w2_full, gg & 1, tw[4])) return false;
                tcgen05_fence_after();
                const uint32_t d = tmem_base + kYCol;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t koff = k * 32;
                    const uint64_t da0 = smem_desc_kmajor_sw128(h1_addr + koff);
                    const uint64_t db0 = smem_desc_kmajor_sw128(w2_addr + koff);
                    umma_f16_pair(d, da0, db0, idesc2, (cc > 0 || k > 0) ? 1u : 0u);
                    if (two) {
                        umma_f16_pair(d, da0, smem_desc_kmajor_sw128(w2_addr + kTileBytes + koff), idesc2, 1);
                        umma_f16_pair(d, smem_desc_kmajor_sw128(h1_addr + kTileBytes + koff), db0, idesc2, 1);
                    }
                }
                umma_commit_pair(&h1_empty);
                umma_commit_pair(&w2_empty);
                return true;
            };
            for (int pt = worker; pt < pair_tiles && ok; pt += workers, ++it) {
                if (!timed_wait(&x_full, (uint32_t)(it & 1), tw[0])) { ok = false; break; }
                for (int c = 0; c < NC && ok; ++c) {
                    const uint32_t g = (uint32_t)(it * NC + c);
                    const int hb = (int)(g & 1);
                    if (!timed_wait(&hacc_empty[hb], ((g >> 1) & 1) ^ 1, tw[1])) { ok = false; break; }
                    const uint32_t d = tmem_base + kHaccCol + hb * kFC;
                    for (int kb = 0; kb < 4 && ok; ++kb) {
                        if (!timed_wait(&w1_full[kb], g & 1, tw[2])) { ok = false; break; }
                        tcgen05_fence_after();
                        const uint32_t a0 = x_addr + kb * 2 * kTileBytes;
                        const uint32_t b0 = w1_addr + kb * kW1SubBytes;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t koff = k * 32;
                            const uint64_t da0 = smem_desc_kmajor_sw128(a0 + koff);
                            const uint64_t db0 = smem_desc_kmajor_sw128(b0 + koff);
                            umma_f16_pair(d, da0, db0, idesc1, (kb > 0 || k > 0) ? 1u : 0u);
                            if (two) {
                                umma_f16_pair(d, da0, smem_desc_kmajor_sw128(b0 + 32 * 128 + koff), idesc1, 1);
                                umma_f16_pair(d, smem_desc_kmajor_sw128(a0 + kTileBytes + koff), db0, idesc1, 1);
                            }
                        }
                        umma_commit_pair(&w1_empty[kb]);
                    }
                    if (!ok) break;
                    umma_commit_pair(&hacc_full[hb]);
                    if (c >= 1 && !issue_g2(g - 1, c - 1)) { ok = false; break; }
                }
                if (ok && !issue_g2((uint32_t)(it * NC + NC - 1), NC - 1)) ok = false;
                if (ok) umma_commit_pair(&y_full);
            }
            if (!ok) atomicExch(p.status, kStatusMmaTimeout);
            if (p.trace) {
                for (int i = 0; i < 6; ++i) atomicAdd(p.trace + i, (unsigned long long)tw[i]);
                atomicAdd(p.trace + 6, (unsigned long long)(clock64() - t_begin));
            }
        }
    } else {
        // ------------------------------------------------ epilogue warps 2..9
        const int set = (warp - 2) >> 2;          // which 32 of a chunk's 64 columns / column half of Y
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const bool elected = (warp - 2) == 4 * set && lane == 0;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t h1_addr = smem_u32(h1_smem);
        const uint32_t hacc_empty_remote[2] = {map_to_cta(&hacc_empty[0], 0), map_to_cta(&hacc_empty[1], 0)};
        const uint32_t h1_full_remote = map_to_cta(&h1_full, 0);
        const uint32_t y_empty_remote = map_to_cta(&y_empty, 0);
        const float scale1 = *p.scale1, scale2 = *p.scale2;
        uint32_t raw[32], h[32], l[32];
        float y[32];
        bool ok = true;
        int it = 0;
        long long te[4] = {0, 0, 0, 0};
        const long long t_begin = clock64();
        for (int pt = worker; pt < pair_tiles && ok; pt += workers, ++it) {
            const int m_blk = 2 * pt + (int)rank;
            const int m0 = m_blk * kBM;
            const int64_t m = (int64_t)m0 + row;
            // ---- hidden chunks: Hacc -> relu -> split -> H1 (A operand of G2).  Epilogue set s
            // (4 warps, one row per thread) owns the chunks with c % 2 == s, i.e. always the
            // accumulator buffer Hacc[s]; the per-chunk synchronisation cost is paid by 4 warps
            // every other chunk instead of 8 warps every chunk.
            float bias2nd[32];
#pragma unroll 1
            for (int c = set; c < NC; c += 2) {
                const uint32_t g = (uint32_t)(it * NC + c);
                // the bias loads (L2 latency) are in flight while this set waits for its chunk
                load_row32(p.bias1 + c * kFC, y);
                load_row32(p.bias1 + c * kFC + 32, bias2nd);
                if (!timed_wait(&hacc_full[set], (g >> 1) & 1, te[0])) { ok = false; break; }
                tcgen05_fence_after();
                uint32_t raw2[32];
                tmem_ld_32x32(t_lane + kHaccCol + set * kFC, raw);
                tmem_ld_32x32(t_lane + kHaccCol + set * kFC + 32, raw2);
                tmem_wait_ld();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(hacc_empty_remote[set]);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float v0 = fmaxf(fmaf(__uint_as_float(raw[2 * j]), scale1, y[2 * j]), 0.f);
                    const float v1 = fmaxf(fmaf(__uint_as_float(raw[2 * j + 1]), scale1, y[2 * j + 1]), 0.f);
                    split2_f16(v0, v1, h[j], l[j]);
                    const float v2 = fmaxf(fmaf(__uint_as_float(raw2[2 * j]), scale1, bias2nd[2 * j]), 0.f);
                    const float v3 = fmaxf(fmaf(__uint_as_float(raw2[2 * j + 1]), scale1, bias2nd[2 * j + 1]), 0.f);
                    split2_f16(v2, v3, h[16 + j], l[16 + j]);
                }
                if (!timed_wait(&h1_empty, (g & 1) ^ 1, te[1])) { ok = false; break; }
                stage_row128(h1_addr, row, h);
                stage_row128(h1_addr + kTileBytes, row, l);
                fence_proxy_async_smem();
                named_bar_sync(1 + set, 128);
                if (elected) mbar_arrive_cluster(h1_full_remote);
            }
            if (!ok) break;
            // ---- LayerNorm epilogue on Y: set s owns columns [128 s, 128 s + 128)
            if (!timed_wait(&y_full, (uint32_t)(it & 1), te[2])) { ok = false; break; }
            tcgen05_fence_after();
            const long long t_ln = clock64();
            const uint32_t t_acc = t_lane + kYCol;
            const SeqInfo s = p.seqs[p.tile_seq[m_blk]];
            const bool in_tensor = (int)(m - s.row0) < s.tensor_len;
            const int c_first = set * 4;
            float sum = 0.f;
            // the residual IS the X tile still resident in shared memory (128B-swizzled
            // [k-block][hi|lo][128 rows][64]): no global re-read
            const uint32_t x_row = smem_u32(x_smem) + (uint32_t)row * 128;
            const uint32_t xsw = (uint32_t)row & 7;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = c_first + i;
                tmem_ld_32x32(t_acc + c * 32, raw);
                load_row32(p.bias2 + c * 32, y);
                uint4 rh[4], rl[4];
                const uint32_t tile = x_row + (uint32_t)(c >> 1) * 2 * kTileBytes;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t addr = tile + (((uint32_t)((c & 1) * 4 + u) ^ xsw) << 4);
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(rh[u].x), "=r"(rh[u].y), "=r"(rh[u].z), "=r"(rh[u].w)
                                 : "r"(addr));
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(rl[u].x), "=r"(rl[u].y), "=r"(rl[u].z), "=r"(rl[u].w)
                                 : "r"(addr + kTileBytes));
                }
                tmem_wait_ld();
                const uint32_t* rhw = reinterpret_cast<const uint32_t*>(rh);
                const uint32_t* rlw = reinterpret_cast<const uint32_t*>(rl);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&rhw[j]));
                    const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&rlw[j]));
                    const float v0 = fmaf(__uint_as_float(raw[2 * j]), scale2, y[2 * j]) + (fh.x + fl.x);
                    const float v1 = fmaf(__uint_as_float(raw[2 * j + 1]), scale2, y[2 * j + 1]) + (fh.y + fl.y);
                    sum += v0 + v1;
                    raw[2 * j] = __float_as_uint(v0);
                    raw[2 * j + 1] = __float_as_uint(v1);
                }
                tmem_st_32x32(t_acc + c * 32, raw);
            }
            // X is dead now (all G1 MMAs completed before y_full): let the producer refill it
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_empty);
            tmem_wait_st();
            ln_part[set][row] = sum;
            named_bar_sync(3, 256);
            const float mean = (ln_part[0][row] + ln_part[1][row]) * (1.f / 256);
            named_bar_sync(3, 256);
            float sq = 0.f;
#pragma unroll 1
            for (int c = c_first; c < c_first + 4; ++c) {
                tmem_ld_32x32(t_acc + c * 32, raw);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float d = __uint_as_float(raw[j]) - mean;
                    sq = fmaf(d, d, sq);
                }
            }
            ln_part[set][row] = sq;
            named_bar_sync(3, 256);
            const float rstd = rsqrtf((ln_part[0][row] + ln_part[1][row]) * (1.f / 256) + p.eps);
            named_bar_sync(3, 256);
            unsigned char* stage = h1_smem + set * kTileBytes;   // H1 is dead until the next tile
#pragma unroll 1
            for (int gq = 0; gq < 2; ++gq) {
                const int n0 = (c_first + 2 * gq) * 32;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float gamma[32];
                    tmem_ld_32x32(t_acc + n0 + half * 32, raw);
                    load_row32(p.gamma + n0 + half * 32, gamma);
                    load_row32(p.beta + n0 + half * 32, y);
                    tmem_wait_ld();
                    if (gq == 1 && half == 1) {   // Y has been read: the next tile may overwrite it
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster_relaxed(y_empty_remote);
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float v0 = fmaf((__uint_as_float(raw[2 * j]) - mean) * rstd, gamma[2 * j], y[2 * j]);
                        const float v1 = fmaf((__uint_as_float(raw[2 * j + 1]) - mean) * rstd, gamma[2 * j + 1], y[2 * j + 1]);
                        split2_f16(in_tensor ? v0 : 0.f, in_tensor ? v1 : 0.f, h[half * 16 + j], l[half * 16 + j]);
                    }
                }
#pragma unroll
                for (int plane = 0; plane < 2; ++plane) {
                    if (elected) bulk_wait_read_all();
                    named_bar_sync(1 + set, 128);
                    stage_row128(smem_u32(stage), row, plane == 0 ? h : l);
                    fence_proxy_async_smem();
                    named_bar_sync(1 + set, 128);
                    if (elected) {
                        tma_store_3d(&map_out, stage, n0, m0, plane);
                        bulk_commit_group();
                    }
                }
            }
            // the staging tiles alias H1: drain the stores before the next tile's chunks
            if (elected) bulk_wait_read_all();
            named_bar_sync(3, 256);
            te[3] += clock64() - t_ln;
        }
        if (elected) bulk_wait_all();
        if (p.trace && warp == 2 && lane == 0) {
            for (int i = 0; i < 4; ++i) atomicAdd(p.trace + 8 + i, (unsigned long long)te[i]);
            atomicAdd(p.trace + 12, (unsigned long long)(clock64() - t_begin));
            atomicAdd(p.trace + 13, 1ull);
        }
        if (!ok) atomicExch(p.status, kStatusEpilogueTimeout);
    }

    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_pair<512>(tmem_base);
    }
}

int launch_ffn_fused(ppgs_engine* e, const CUtensorMap& map_x, const CUtensorMap& map_w1,
                     const CUtensorMap& map_w2, const CUtensorMap& map_out, const FfnParams& p,
                     cudaStream_t stream) {
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
    cudaLaunchAttribute attrs[1];
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
    {
        LaunchScope scope(e, "tc_ffn_fused_ln", stream);
        PPGS_CUDA(cudaLaunchKernelEx(&cfg, ffn_fused_kernel, map_x, map_w1, map_w2, map_out, p));
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

}  // namespace tc
}  // namespace ppgs
