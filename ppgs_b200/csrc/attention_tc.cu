// Length-masked multi-head attention of the tensor-core path.
//
// Replaces F.scaled_dot_product_attention inside nn.MultiheadAttention as called
// from ppgs/model/transformer.py:76-80 (key-padding mask from `lengths`, optional
// square subsequent mask when IS_CAUSAL).
#include <float.h>

#include "attention_tc.cuh"
#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace ppgs {

using namespace tc;

// ---------------------------------------------------------------------------
// CUDA-core kernel over the split planes (fp32 arithmetic): block = (32 queries,
// head, sequence), warp = 4 queries, lane = key inside a 32-key tile.
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
attention_planes_kernel(const __half* __restrict__ qkv, int64_t plane_stride, int H,
                        const SeqInfo* __restrict__ seqs, int causal, float scale, int planes,
                        __half* __restrict__ out, int64_t out_plane_stride) {
    constexpr int QT = 32, KT = 32, DP = D + 1, PER = D / 32;
    extern __shared__ float smem[];
    float* qs = smem;
    float* ks = qs + QT * D;
    float* vs = ks + KT * DP;
    const SeqInfo s = seqs[blockIdx.z];
    const int q0 = blockIdx.x * QT;
    if (q0 >= s.tensor_len) return;
    const int head = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t ld = 3 * H;
    const __half* base = qkv + (int64_t)s.row0 * ld;
    auto load = [&](const __half* p) {
        float v = __half2float(*p);
        if (planes == 2) v += __half2float(p[plane_stride]);
        return v;
    };
    for (int i = tid; i < QT * D; i += 256) {
        int q = i / D, d = i - q * D;
        qs[i] = load(base + (int64_t)(q0 + q) * ld + head * D + d);
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
            const __half* r = base + (int64_t)(k0 + j) * ld + head * D + d;
            ks[j * DP + d] = load(r + H);
            vs[j * D + d] = load(r + 2 * H);
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
            const float corr = expf(m[q] - mnew);
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
        const float inv = l[q] > 0.f ? 1.f / l[q] : 0.f;
        __half* dst = out + (int64_t)(s.row0 + t) * H + head * D;
#pragma unroll
        for (int c = 0; c < PER; ++c) {
            __half hi, lo;
            split_f16(acc[q][c] * inv, hi, lo);
            dst[lane + 32 * c] = hi;
            dst[out_plane_stride + lane + 32 * c] = lo;
        }
    }
}


// ---------------------------------------------------------------------------
// tcgen05 kernel: one CTA per (128-query tile, head, sequence), keys <= 512,
// head_dim D in {64, 128, 256}.
//
//   warp 0      TMA producer: Q tile, K blocks (KB keys) through a ring, then V chunks
//               (64 keys) through a ring that reuses the K ring
//   warp 1      tcgen05.mma issuer: S_j = Q K_j^T into TMEM columns [KB j, KB j + KB),
//               then O += P_c V_c into the last D columns
//   warps 2-9   two threads per query row: row max over S (TMEM), p = exp2((s - max) c),
//               split-fp16 P chunks written to the 128B-swizzled smem layout the MMA
//               reads (ring that reuses the Q tile), final O / sum -> planes
//
// S for 512 keys fills all 512 TMEM columns; O reuses the last D of them, so the 64-key
// chunks that live there are converted to P first (chunk_order) and the first P.V MMA
// waits for all of them — the P ring has at least that many slots.
//
// Shared memory: region A = Q tile [D/64 d chunks][2 planes][128 x 64], later the P ring
// and the output staging tiles; region B = K ring, later the V ring.  D = 256 only fits
// one K / V slot beside its 128 KB Q tile (loads and MMAs alternate there).
// ---------------------------------------------------------------------------
constexpr int kAttnThreads = 320;   // producer + MMA + 8 softmax warps
constexpr int kTile16K = 16384;                // [128 rows][64 fp16]
constexpr int kPSlotBytes = 2 * kTile16K;      // [plane 2][128 q][64 keys]
constexpr int kVPlaneBytes = 8192;             // [64 keys][64 d]

template <int D>
struct AttnShape {
    static constexpr int kDC = D / 64;                          // 64-wide d chunks
    static constexpr int kKB = D == 256 ? 64 : 128;             // keys per K block
    static constexpr int kKTile = kKB * 128;                    // [KB keys][64 fp16]
    static constexpr int kQBytes = kDC * 2 * kTile16K;
    static constexpr int kNP = D == 256 ? 4 : 2;                // P ring slots
    static constexpr int kABytes = kQBytes > kNP * kPSlotBytes ? kQBytes : kNP * kPSlotBytes;
    static constexpr int kKSlotBytes = kDC * 2 * kKTile;
    static constexpr int kNK = D == 256 ? 1 : 2;
    static constexpr int kVSlotBytes = kDC * 2 * kVPlaneBytes;
    static constexpr int kNV = kNK * kKSlotBytes / kVSlotBytes;
    static constexpr int kOCol = 512 - D;
    static constexpr int kSafeChunks = kOCol / 64;              // chunks whose S columns O never touches
    static constexpr size_t kSmem = kABytes + kNK * kKSlotBytes + 1024;
    static_assert(8 - kSafeChunks <= kNP, "P ring must hold every chunk that overlaps O");
    static_assert(kNV >= 1 && kNV <= 4 && kNK <= 2 && kNP <= 4, "ring sizes");
};

struct AttnParams {
    const SeqInfo* seqs;
    int H, causal, planes;   // planes: V operand and output planes
    int qk_planes, p_planes; // Q / K and P operands: 2 = hi + lo, 1 = hi only (fewer MMA passes)
    float scale_log2e;
    __half* out;
    int64_t out_plane_stride;
    int q_first_tile;   // streaming decoder: only query tiles >= this one are computed
    int q_first_per_seq;   // 1: per sequence, from SeqInfo::src_start
    int* status;
    // cycle accounting (PPGS_B200_TRACE=1): MMA [0] wait q [1] wait k [2] wait p [3] wait v [4] total;
    // softmax warp 2: [8] wait s [9] row max [10] wait p_empty [11] chunk work [12] wait o [13] total [14] CTAs
    unsigned long long* trace;
};

// 2^x for x <= 0 (softmax exponents): one MUFU.EX2, flushes to zero on underflow
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int SAFE>
__device__ __forceinline__ int chunk_order(int i, int nchunks) {
    if (nchunks <= SAFE) return i;
    const int pre = nchunks - SAFE;
    return i < pre ? SAFE + i : i - pre;
}


template <int D>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_out,
                    const AttnParams p) {
    using Shape = AttnShape<D>;
    constexpr int DC = Shape::kDC, KB = Shape::kKB, NP = Shape::kNP, NK = Shape::kNK, NV = Shape::kNV;
    constexpr int SAFE = Shape::kSafeChunks;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* q_smem = smem;                      // later: P ring, output staging
    unsigned char* k_ring = smem + Shape::kABytes;     // later: V ring
    __shared__ __align__(8) uint64_t q_full, o_full, s_full[8];   // s_full[j]: scores of key block j
    __shared__ __align__(8) uint64_t k_full[2], k_empty[2], p_full[4], p_empty[4];
    __shared__ __align__(8) uint64_t v_full[4], v_empty[4];
    __shared__ uint32_t tmem_slot;
    __shared__ float row_part[2][128];   // per-row partial max / sum of the two halves

    const SeqInfo s = p.seqs[blockIdx.z];
    const int q0 = ((int)blockIdx.x + (p.q_first_per_seq ? s.src_start : p.q_first_tile)) * 128, head = blockIdx.y;
    if (q0 >= s.tensor_len) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int nkeys = s.valid_len;
    if (p.causal) nkeys = min(nkeys, q0 + 128);
    const int64_t out_row0 = (int64_t)(s.row0 + q0);
    pdl_launch_dependents();
    if (nkeys <= 0) {
        // every key masked (chunk_lengths == 0 rows of transformer.py:59-60): zeros
        pdl_wait();
        for (int i = threadIdx.x; i < 128 * D; i += kAttnThreads) {
            const int r = i / D, d = i - r * D;
            __half* dst = p.out + (out_row0 + r) * p.H + head * D + d;
            dst[0] = __float2half_rn(0.f);
            dst[p.out_plane_stride] = __float2half_rn(0.f);
        }
        return;
    }
    const int nb = (nkeys + KB - 1) / KB, nchunks = (nkeys + 63) >> 6;

    if (threadIdx.x == 0) {
        mbar_init(&q_full, 1);
        mbar_init(&o_full, 1);
        for (int i = 0; i < 8; ++i) mbar_init(&s_full[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&p_full[i], 8);
            mbar_init(&p_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(&tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_wait();   // Q / K / V of the previous kernel are needed from here on

    if (warp == 0) {
        if (lane == 0) {
            prefetch_tensormap(&map_q);
            prefetch_tensormap(&map_k);
            prefetch_tensormap(&map_v);
            const int col_q = head * D, col_k = p.H + head * D, col_v = 2 * p.H + head * D;
            bool ok = true;
            mbar_arrive_expect_tx(&q_full, p.qk_planes * DC * kTile16K);
            for (int dc = 0; dc < DC; ++dc)
                tma_load_3d(q_smem + dc * 2 * kTile16K, &map_q, &q_full, col_q + dc * 64, s.row0 + q0, 0);
            for (int j = 0; j < nb && ok; ++j) {
                const int slot = j % NK;
                if (!mbar_wait(&k_empty[slot], ((j / NK) & 1) ^ 1)) { ok = false; break; }
                mbar_arrive_expect_tx(&k_full[slot], p.qk_planes * DC * Shape::kKTile);
                for (int dc = 0; dc < DC; ++dc)
                    tma_load_3d(k_ring + slot * Shape::kKSlotBytes + dc * 2 * Shape::kKTile, &map_k,
                                &k_full[slot], col_k + dc * 64, s.row0 + j * KB, 0);
            }
            if (ok && !mbar_wait(&s_full[nb - 1], 0)) ok = false;   // K ring is dead: reuse it for V
            for (int i = 0; i < nchunks && ok; ++i) {
                const int c = chunk_order<SAFE>(i, nchunks), slot = i % NV;
                if (!mbar_wait(&v_empty[slot], ((i / NV) & 1) ^ 1)) { ok = false; break; }
                mbar_arrive_expect_tx(&v_full[slot], p.planes * DC * kVPlaneBytes);
                for (int dh = 0; dh < DC; ++dh)
                    tma_load_3d(k_ring + slot * Shape::kVSlotBytes + dh * 2 * kVPlaneBytes, &map_v,
                                &v_full[slot], col_v + dh * 64, s.row0 + c * 64, 0);
            }
            if (!ok) atomicExch(p.status, kStatusAttnTimeout);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_f16(128, KB, false);
            constexpr uint32_t idesc_o = make_idesc_f16(128, D, true);
            const uint32_t q_addr = smem_u32(q_smem), k_addr = smem_u32(k_ring);
            long long tw[4] = {0, 0, 0, 0};
            const long long t_begin = clock64();
            auto timed = [&](uint64_t* bar, uint32_t parity, int slot_id) {
                const long long t0 = clock64();
                const bool got = mbar_wait(bar, parity);
                tw[slot_id] += clock64() - t0;
                return got;
            };
            bool ok = timed(&q_full, 0, 0);
            tcgen05_fence_after();
            for (int j = 0; j < nb && ok; ++j) {
                const int slot = j % NK;
                if (!timed(&k_full[slot], (j / NK) & 1, 1)) { ok = false; break; }
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + j * KB;
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks) {
                    const uint32_t qa = q_addr + (ks >> 2) * 2 * kTile16K + (ks & 3) * 32;
                    const uint32_t kb = k_addr + slot * Shape::kKSlotBytes + (ks >> 2) * 2 * Shape::kKTile +
                                        (ks & 3) * 32;
                    const uint64_t dq0 = smem_desc_kmajor_sw128(qa), dk0 = smem_desc_kmajor_sw128(kb);
                    umma_f16(d_tmem, dq0, dk0, idesc_s, ks > 0);
                    if (p.qk_planes == 2) {
                        umma_f16(d_tmem, dq0, smem_desc_kmajor_sw128(kb + Shape::kKTile), idesc_s, 1);
                        umma_f16(d_tmem, smem_desc_kmajor_sw128(qa + kTile16K), dk0, idesc_s, 1);
                    }
                }
                umma_commit(&k_empty[slot]);
                umma_commit(&s_full[j]);   // the softmax warps start their row-max pass on block j
            }
            const uint32_t o_tmem = tmem_base + Shape::kOCol;
            const int pre = nchunks > SAFE ? nchunks - SAFE : 0;
            for (int i = 0; i < nchunks && ok; ++i) {
                const int ps = i % NP, vs = i % NV;
                if (!timed(&p_full[ps], (i / NP) & 1, 2)) { ok = false; break; }
                if (i == 0)   // every chunk whose scores live in O's columns has been converted
                    for (int k = 1; k < pre && ok; ++k) ok = timed(&p_full[k], 0, 2);
                if (!ok) break;
                if (!timed(&v_full[vs], (i / NV) & 1, 3)) { ok = false; break; }
                tcgen05_fence_after();
                const uint32_t p_addr = q_addr + ps * kPSlotBytes;
                const uint32_t v_addr = k_addr + vs * Shape::kVSlotBytes;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t pa = p_addr + ks * 32;          // 16 keys along the swizzle row
                    const uint32_t vb = v_addr + ks * 16 * 128;    // 16 key rows
                    const uint64_t dp0 = smem_desc_kmajor_sw128(pa);
                    const uint64_t dv0 = smem_desc_mnmajor_sw128(vb, 2 * kVPlaneBytes);
                    umma_f16(o_tmem, dp0, dv0, idesc_o, (i > 0 || ks > 0) ? 1u : 0u);
                    if (p.planes == 2)
                        umma_f16(o_tmem, dp0, smem_desc_mnmajor_sw128(vb + kVPlaneBytes, 2 * kVPlaneBytes),
                                 idesc_o, 1);
                    if (p.p_planes == 2)
                        umma_f16(o_tmem, smem_desc_kmajor_sw128(pa + kTile16K), dv0, idesc_o, 1);
                }
                umma_commit(&p_empty[ps]);
                umma_commit(&v_empty[vs]);
            }
            if (ok) umma_commit(&o_full);
            else atomicExch(p.status, kStatusAttnTimeout);
            if (p.trace) {
                for (int i = 0; i < 4; ++i) atomicAdd(p.trace + i, (unsigned long long)tw[i]);
                atomicAdd(p.trace + 4, (unsigned long long)(clock64() - t_begin));
            }
        }
    } else {
        // 8 softmax warps: a query row is shared by two threads (same TMEM lane
        // quadrant, `half` = which 32 of every 64 keys / which D/2 of the D output
        // columns); row max and row sum are combined through shared memory.
        const int half = (warp - 2) >> 2;
        const int quad = warp & 3, r = quad * 32 + lane, t = q0 + r;
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
        uint32_t raw[32];
        const long long ts0 = clock64();
        bool ok = true;
        long long t_pempty = 0, ts1 = ts0;
        float mx = -FLT_MAX;
        {
            // row max, block by block as the S MMAs retire (the first wait is the exposed one)
            const int ngroups = (nkeys + 31) >> 5;
            int ready = -1;
#pragma unroll 1
            for (int g = half; g < ngroups && ok; g += 2) {
                const int blk = (g * 32) / KB;
                if (blk > ready) {
                    ok = mbar_wait(&s_full[blk], 0);
                    if (ready < 0) ts1 = clock64();
                    ready = blk;
                    tcgen05_fence_after();
                    if (!ok) break;
                }
                tmem_ld_32x32(t_row + g * 32, raw);
                tmem_wait_ld();
                const bool full = g * 32 + 32 <= nkeys && (!p.causal || g * 32 + 31 <= q0);
                if (full) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(raw[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int key = g * 32 + j;
                        const bool allowed = key < nkeys && (!p.causal || key <= t);
                        mx = fmaxf(mx, allowed ? __uint_as_float(raw[j]) : -FLT_MAX);
                    }
                }
            }
        }
        // every thread has observed the completion of ALL score MMAs before the exp pass
        // (a thread whose groups ended early never waited for the last block)
        if (ok) ok = mbar_wait(&s_full[nb - 1], 0);
        tcgen05_fence_after();
        row_part[half][r] = mx;
        named_bar_sync(1, 256);
        mx = fmaxf(row_part[0][r], row_part[1][r]);
        named_bar_sync(1, 256);   // row_part is reused for the row sums
        const long long ts2 = clock64();
        const float mc = mx * p.scale_log2e;
        float sum = 0.f;
        const uint32_t p_base = smem_u32(q_smem);
        const uint32_t row_off = (uint32_t)r * 128, sw = (uint32_t)(r & 7);
        // one chunk: `cur` holds this thread's 32 scores (its tcgen05.ld was issued one
        // chunk earlier); the next chunk's load is issued before the math so TMEM latency
        // hides behind it, and the P slot is only waited for right before it is written
        auto chunk = [&](int i, uint32_t (&cur)[32], uint32_t (&nxt)[32]) -> bool {
            const int c = chunk_order<SAFE>(i, nchunks), slot = i % NP;
            const int key0 = c * 64 + half * 32;
            tmem_wait_ld();
            if (i + 1 < nchunks)
                tmem_ld_32x32(t_row + chunk_order<SAFE>(i + 1, nchunks) * 64 + half * 32, nxt);
            uint32_t h[16], l[16];
            const bool full = key0 + 32 <= nkeys && (!p.causal || key0 + 31 <= q0);
            if (full && p.p_planes == 1) {
                // single-plane P: the row sum is taken over the ROUNDED numerators, so the
                // weights the P.V MMA applies sum to one exactly
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float e0 = fast_exp2(fmaf(__uint_as_float(cur[2 * j]), p.scale_log2e, -mc));
                    const float e1 = fast_exp2(fmaf(__uint_as_float(cur[2 * j + 1]), p.scale_log2e, -mc));
                    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(e1), "f"(e0));
                    const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
                    sum += back.x + back.y;
                }
            } else if (full) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float e0 = fast_exp2(fmaf(__uint_as_float(cur[2 * j]), p.scale_log2e, -mc));
                    const float e1 = fast_exp2(fmaf(__uint_as_float(cur[2 * j + 1]), p.scale_log2e, -mc));
                    sum += e0 + e1;
                    split2_f16(e0, e1, h[j], l[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float pv[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int key = key0 + 2 * j + q;
                        const bool allowed = key < nkeys && (!p.causal || key <= t);
                        const float e = fast_exp2(fmaf(__uint_as_float(cur[2 * j + q]), p.scale_log2e, -mc));
                        pv[q] = allowed ? e : 0.f;
                    }
                    if (p.p_planes == 1) {
                        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(pv[1]), "f"(pv[0]));
                        const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
                        sum += back.x + back.y;
                    } else {
                        sum += pv[0] + pv[1];
                        split2_f16(pv[0], pv[1], h[j], l[j]);
                    }
                }
            }
            {
                const long long t0 = clock64();
                const bool got = mbar_wait(&p_empty[slot], ((i / NP) & 1) ^ 1);
                t_pempty += clock64() - t0;
                if (!got) return false;
            }
            const uint32_t p_slot = p_base + slot * kPSlotBytes;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t unit = (uint32_t)(half * 4 + u) ^ sw;
                const uint32_t addr = p_slot + row_off + unit * 16;
                st_shared_v4(addr, h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]);
                if (p.p_planes == 2)
                    st_shared_v4(addr + kTile16K, l[4 * u], l[4 * u + 1], l[4 * u + 2], l[4 * u + 3]);
            }
            fence_proxy_async_smem();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[slot]);
            return true;
        };
        uint32_t raw_b[32];
        if (ok) tmem_ld_32x32(t_row + chunk_order<SAFE>(0, nchunks) * 64 + half * 32, raw);
#pragma unroll 1
        for (int i = 0; i < nchunks && ok; i += 2) {
            ok = chunk(i, raw, raw_b);
            if (ok && i + 1 < nchunks) ok = chunk(i + 1, raw_b, raw);
        }
        tmem_wait_ld();
        row_part[half][r] = sum;
        named_bar_sync(1, 256);
        sum = row_part[0][r] + row_part[1][r];
        const long long ts3 = clock64();
        if (ok && !mbar_wait(&o_full, 0)) ok = false;
        const long long ts4 = clock64();
        tcgen05_fence_after();
        if (ok) {
            // O / sum -> split planes, staged as [128 rows][64 cols] 128B-swizzled tiles
            // (64-column block x plane) in the dead P ring, then stored with TMA: full
            // 128-byte rows instead of 16-byte pieces per thread
            const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
            for (int g = 0; g < D / 64; ++g) {
                const int col = half * (D / 2) + g * 32;   // this thread's 32 output columns
                tmem_ld_32x32(t_row + Shape::kOCol + col, raw);
                tmem_wait_ld();
                uint32_t h[16], l[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    split2_f16(__uint_as_float(raw[2 * j]) * inv, __uint_as_float(raw[2 * j + 1]) * inv,
                               h[j], l[j]);
                const uint32_t tile_hi = p_base + (uint32_t)((col >> 6) * 2) * kTile16K;
                const uint32_t unit0 = (uint32_t)(col & 63) >> 3;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t addr = tile_hi + row_off + (((unit0 + u) ^ sw) << 4);
                    st_shared_v4(addr, h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]);
                    st_shared_v4(addr + kTile16K, l[4 * u], l[4 * u + 1], l[4 * u + 2], l[4 * u + 3]);
                }
            }
            fence_proxy_async_smem();
        }
        named_bar_sync(1, 256);
        if (ok && warp == 2 && lane == 0) {
#pragma unroll
            for (int i = 0; i < 2 * DC; ++i)
                tma_store_3d(&map_out, q_smem + i * kTile16K, head * D + (i >> 1) * 64,
                             (int)out_row0, i & 1);
            bulk_commit_group();
            bulk_wait_all();
        }
        if (!ok) atomicExch(p.status, kStatusAttnTimeout);
        if (p.trace && warp == 2 && lane == 0) {
            atomicAdd(p.trace + 8, (unsigned long long)(ts1 - ts0));
            atomicAdd(p.trace + 9, (unsigned long long)(ts2 - ts1));
            atomicAdd(p.trace + 10, (unsigned long long)t_pempty);
            atomicAdd(p.trace + 11, (unsigned long long)(ts3 - ts2 - t_pempty));
            atomicAdd(p.trace + 12, (unsigned long long)(ts4 - ts3));
            atomicAdd(p.trace + 13, (unsigned long long)(clock64() - ts0));
            atomicAdd(p.trace + 14, 1ull);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

int launch_attention_planes(ppgs_engine* e, int head_dim, const __half* qkv, __half* out, int rows,
                            int H, int heads, int max_pitch, int nseq, const SeqInfo* seqs_dev,
                            int causal, int planes, cudaStream_t stream) {
    auto run = [&](auto kernel, int D) -> int {
        const size_t smem = (size_t)(32 * D + 32 * (D + 1) + 32 * D) * sizeof(float);
        PPGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(max_pitch / 32, heads, (unsigned)nseq);
        {
            LaunchScope scope(e, "attention_simt_planes", stream);
            kernel<<<grid, 256, smem, stream>>>(qkv, (int64_t)rows * 3 * H, H, seqs_dev, causal,
                                                1.f / sqrtf((float)D), planes, out, (int64_t)rows * H);
        }
        PPGS_CUDA(cudaGetLastError());
        return PPGS_OK;
    };
    if (head_dim == 64) return run(attention_planes_kernel<64>, 64);
    if (head_dim == 128) return run(attention_planes_kernel<128>, 128);
    if (head_dim == 256) return run(attention_planes_kernel<256>, 256);
    set_error("attention_planes: head_dim %d not built", head_dim);
    return PPGS_E_UNSUPPORTED;
}

template <int D>
static int run_attention_tc(ppgs_engine* e, const __half* qkv, __half* out, int rows, int H, int heads,
                            int max_pitch, int nseq, const SeqInfo* seqs_dev, int causal, int planes,
                            cudaStream_t stream, int q_first_tile, int q_tiles, int qk_planes, int p_planes) {
    using Shape = AttnShape<D>;
    const bool per_seq = q_first_tile < 0;   // -1: first query tile per sequence (SeqInfo::src_start)
    CUtensorMap map_q, map_k, map_v, map_out;
    PPGS_CHECK(make_store_map(&map_out, out, H, rows, (uint64_t)rows * H));
    PPGS_CHECK(make_plane_map(&map_q, qkv, false, 3 * H, rows, 1, 2, 3 * H, 0,
                              (uint64_t)rows * 3 * H, 128, qk_planes));
    PPGS_CHECK(make_plane_map(&map_k, qkv, false, 3 * H, rows, 1, 2, 3 * H, 0,
                              (uint64_t)rows * 3 * H, Shape::kKB, qk_planes));
    PPGS_CHECK(make_plane_map(&map_v, qkv, false, 3 * H, rows, 1, 2, 3 * H, 0,
                              (uint64_t)rows * 3 * H, 64, planes));
    static PerDeviceOnce attr_tc;
    if (attr_tc.first(e->device)) {
        PPGS_CUDA(cudaFuncSetAttribute(attention_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Shape::kSmem));
    }
    AttnParams p;
    p.seqs = seqs_dev;
    p.H = H;
    p.causal = causal;
    p.planes = planes;
    p.qk_planes = qk_planes;
    p.p_planes = p_planes;
    p.scale_log2e = 1.4426950408889634f / sqrtf((float)D);
    p.out = out;
    p.out_plane_stride = (int64_t)rows * H;
    p.status = e->status_dev;
    p.q_first_tile = per_seq ? 0 : q_first_tile;
    p.q_first_per_seq = per_seq ? 1 : 0;
    p.trace = e->trace_dev ? e->trace_dev + 80 : nullptr;
    dim3 grid(q_tiles > 0 ? q_tiles : max_pitch / 128, heads, (unsigned)nseq);
    {
        LaunchScope scope(e, "tc_attention", stream);
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attrs[1];
        cfg.gridDim = grid;
        cfg.blockDim = dim3(kAttnThreads);
        cfg.dynamicSmemBytes = Shape::kSmem;
        cfg.stream = stream;
        if (pdl_enabled()) {
            attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attrs[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attrs;
            cfg.numAttrs = 1;
        }
        PPGS_CUDA(cudaLaunchKernelEx(&cfg, attention_tc_kernel<D>, map_q, map_k, map_v, map_out, p));
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

int launch_attention_any(ppgs_engine* e, const __half* qkv, __half* out, int rows, int H, int heads,
                         int max_pitch, int nseq, const SeqInfo* seqs_dev, int causal, int planes,
                         cudaStream_t stream, int q_first_tile, int q_tiles, int qk_planes, int p_planes) {
    const int D = H / heads;
    qk_planes = (qk_planes == 1 || planes == 1) ? 1 : 2;
    p_planes = (p_planes == 1 || planes == 1) ? 1 : 2;
    // default model (head_dim 128): two query tiles per CTA, K / V streamed in 128-key blocks
    // (any sequence length), single-plane Q / K / P
    if (e->attention_impl == 1 && e->attn_dual && e->status_dev &&
        attention_dual_supported(D, max_pitch, qk_planes, p_planes))
        return launch_attention_dual(e, qkv, out, rows, H, heads, max_pitch, nseq, seqs_dev, causal, planes, stream,
                                     q_first_tile, q_tiles);
    if (e->attention_impl == 1 && max_pitch <= 512 && max_pitch % 128 == 0 && e->status_dev) {
        if (D == 64)
            return run_attention_tc<64>(e, qkv, out, rows, H, heads, max_pitch, nseq, seqs_dev, causal, planes, stream,
                                         q_first_tile, q_tiles, qk_planes, p_planes);
        if (D == 128)
            return run_attention_tc<128>(e, qkv, out, rows, H, heads, max_pitch, nseq, seqs_dev, causal, planes, stream,
                                         q_first_tile, q_tiles, qk_planes, p_planes);
        if (D == 256)
            return run_attention_tc<256>(e, qkv, out, rows, H, heads, max_pitch, nseq, seqs_dev, causal, planes, stream,
                                         q_first_tile, q_tiles, qk_planes, p_planes);
    }
    if (q_first_tile || q_tiles) {
        set_error("attention: the query-tile window needs the tcgen05 kernel (head_dim 64 / 128 / 256, pitch <= 512)");
        return PPGS_E_UNSUPPORTED;
    }
    return launch_attention_planes(e, D, qkv, out, rows, H, heads, max_pitch, nseq, seqs_dev, causal, planes,
                                   stream);
}

int launch_attention_tc(ppgs_engine* e, const __half* qkv, __half* out, int rows,
                        const ForwardPlan& plan, const SeqInfo* seqs_dev, int planes,
                        cudaStream_t stream) {
    return launch_attention_any(e, qkv, out, rows, e->cfg.hidden_channels, e->cfg.num_heads, plan.max_pitch,
                                (int)plan.seqs.size(), seqs_dev, e->cfg.is_causal, planes, stream, 0, 0,
                                e->attn_qk_planes, e->attn_p_planes);
}

}  // namespace ppgs
