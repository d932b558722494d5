// Length-masked multi-head attention, head_dim 128, two query tiles per CTA.
//
// Replaces F.scaled_dot_product_attention inside nn.MultiheadAttention as called from
// ppgs/model/transformer.py:76-80 (key-padding mask from `lengths`, optional square
// subsequent mask when IS_CAUSAL) for the default model (hidden 256, 2 heads).
//
// Why a second kernel: attention_tc_kernel owns all 512 TMEM columns for ONE query tile and
// runs S = Q K^T -> softmax -> P V one after the other, so the tensor pipe idles while the
// softmax warps work and vice versa (round-1 trace: 25 k cycles per tile, 6 k of them MMA).
// Here a CTA runs TWO query tiles of the same (sequence, head) as independent "lanes" that
// share every K / V block (loaded once, used twice) and take turns on the tensor pipe:
//
//   lane L (L = 0, 1):  TMEM columns [256 L, 256 L + 128) = S block (128 queries x 128 keys),
//                                     [256 L + 128, 256 L + 256) = O accumulator
//   per 128-key block j:   S_j = Q K_j^T (one fp16 pass: Q, K enter as their hi planes)
//                          softmax warps (one thread per query row): scores -> registers,
//                            online max with lazy rescale, p = exp2(s c - m c) -> fp16 -> smem
//                          O += P_j V_j  (two passes: P . V_hi + P . V_lo)
//
// The running max only moves when a block's max exceeds it by more than 2^8 (the rescale of
// O through TMEM is then done by the row's own thread); fp16 holds p <= 2^8 exactly as well
// as p <= 1, and the row sum is taken over the ROUNDED numerators, so the result is the
// reference's softmax(QK^T / sqrt(d)) V up to the fp16 rounding of P (1.8e-5 on
// the posteriorgram) — no approximation of the max is visible in the output.
//
// Warp roles (384 threads): 0 TMA producer, 1 / 2 MMA issuer of lane 0 / 1, 3 TMEM
// allocation, 4-7 softmax of lane 0, 8-11 softmax of lane 1 (TMEM lane quadrant = warp % 4).
// Shared memory (224 KB): Q tiles 2 x 32 KB, P tiles 2 x 32 KB, one K block (32 KB), one V
// block (hi + lo, 64 KB); the output staging tiles reuse the lane's dead Q / P tiles.
#include <float.h>

#include "attention_tc.cuh"
#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace ppgs {

using namespace tc;

namespace {

constexpr int kDualThreads = 384;
constexpr int kD = 128;                      // head_dim
constexpr int kKeys = 128;                   // keys per block
constexpr int kTile = 16384;                 // [128 rows][64 fp16], 128-byte swizzle
constexpr int kQOff = 0;                     // [lane][dc 2][128 q][64]
constexpr int kPOff = 2 * 2 * kTile;         // [lane][kc 2][128 q][64 keys]
constexpr int kKOff = kPOff + 2 * 2 * kTile; // [dc 2][128 keys][64]
constexpr int kVOff = kKOff + 2 * kTile;     // [dh 2][plane 2][128 keys][64]
constexpr int kDualSmem = kVOff + 4 * kTile + 1024;
constexpr float kRescaleLog2 = 8.f;          // running max lags the true max by at most 2^8

struct DualParams {
    const SeqInfo* seqs;
    int H, causal, v_planes;
    uint64_t dead_policy;        // L2 policy of the Q / K / V loads: qkv is dead after this kernel
    int reverse;                 // sequences from the last to the first (serpentine order, GemmParams::reverse)
    float scale_log2e;
    __half* out;
    int64_t out_plane_stride;
    int q_first_tile, q_first_per_seq, q_tiles;   // query-tile window (streaming decoder)
    int* status;
    // cycle accounting (PPGS_B200_TRACE=1), sums over CTAs: lane-0 MMA thread waits [0] q [1] s_empty
    // [2] k_full [3] p_full [4] v_full [5] total; lane-0 softmax warp: [8] wait s_full [9] wait o_done
    // [10] loop total [11] epilogue [12] from kernel start to first scores [14] CTAs
    unsigned long long* trace;
};

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

__global__ void __launch_bounds__(kDualThreads, 1)
attention_dual_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_v,
                      const __grid_constant__ CUtensorMap map_out, const DualParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t q_full, k_full, k_empty, v_full, v_empty;
    __shared__ __align__(8) uint64_t s_full[2], s_empty[2], p_full[2], o_done[2];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t_kernel = clock64();
    const SeqInfo s = p.seqs[p.reverse ? gridDim.z - 1 - blockIdx.z : blockIdx.z];
    const int head = blockIdx.y;
    const int first = p.q_first_per_seq ? s.src_start : p.q_first_tile;
    // lane L owns query tile first + 2 x + L; a lane is idle when its tile lies outside the
    // window or the sequence, and "empty" when every key is masked (zeros, like softmax of
    // chunk_lengths == 0 rows never feeding anything: transformer.py:59-60)
    // (scalars + selects, not arrays: a runtime lane index would put them in local memory)
    auto tile_q0 = [&](int L) { return (first + 2 * (int)blockIdx.x + L) * 128; };
    auto tile_active = [&](int L) {
        const int t = 2 * (int)blockIdx.x + L;
        return (p.q_tiles <= 0 || t < p.q_tiles) && tile_q0(L) < s.tensor_len;
    };
    auto tile_nkeys = [&](int L) {
        int n = s.valid_len;
        if (p.causal) n = min(n, tile_q0(L) + 128);
        return tile_active(L) ? max(n, 0) : 0;
    };
    const int q0_0 = tile_q0(0), q0_1 = tile_q0(1);
    const bool active_0 = tile_active(0), active_1 = tile_active(1);
    const int nkeys_0 = tile_nkeys(0), nkeys_1 = tile_nkeys(1);
    const int nb_0 = (nkeys_0 + kKeys - 1) / kKeys, nb_1 = (nkeys_1 + kKeys - 1) / kKeys;
    const int nb_max = max(nb_0, nb_1);
    pdl_launch_dependents();
    if (!active_0 && !active_1) return;

    if (threadIdx.x == 0) {
        mbar_init(&q_full, 1);
        mbar_init(&k_full, 1);
        mbar_init(&k_empty, 2);
        mbar_init(&v_full, 1);
        mbar_init(&v_empty, 2);
        for (int L = 0; L < 2; ++L) {
            mbar_init(&s_full[L], 1);
            mbar_init(&s_empty[L], 4);
            mbar_init(&p_full[L], 4);
            mbar_init(&o_done[L], 1);
        }
        fence_barrier_init();
    }
    __syncthreads();   // every barrier is initialised before any thread (or the async proxy) touches one
    if (threadIdx.x == 0) {
        // the Q tiles and the first K block are requested before TMEM allocation and the second
        // CTA barrier: their latency (2 us from a cold L2) is the longest item of the prologue
        if (nb_max > 0) {
            pdl_wait();   // Q / K / V of the previous kernel are needed from here on
            const int col_q = head * kD, col_k = p.H + head * kD;
            mbar_arrive_expect_tx(&q_full, ((nb_0 > 0) + (nb_1 > 0)) * 2 * kTile);
            for (int dc = 0; dc < 2; ++dc) {
                if (nb_0 > 0)
                    tma_load_3d_hint(smem + kQOff + dc * kTile, &map_qk, &q_full, col_q + dc * 64, s.row0 + q0_0, 0, p.dead_policy);
                if (nb_1 > 0)
                    tma_load_3d_hint(smem + kQOff + (2 + dc) * kTile, &map_qk, &q_full, col_q + dc * 64, s.row0 + q0_1, 0, p.dead_policy);
            }
            mbar_arrive_expect_tx(&k_full, 2 * kTile);
            for (int dc = 0; dc < 2; ++dc)
                tma_load_3d_hint(smem + kKOff + dc * kTile, &map_qk, &k_full, col_k + dc * 64, s.row0, 0, p.dead_policy);
        }
    }
    if (warp == 3) tmem_alloc<512>(&tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_wait();

    // register budget: the producer / MMA warpgroup hands its registers to the two softmax
    // warpgroups (128 scores + 64 packed numerators live per thread)
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer: Q tiles, K blocks
        if (lane == 0 && nb_max > 0) {
            const int col_k = p.H + head * kD;
            bool ok = true;
            for (int j = 1; j < nb_max && ok; ++j) {   // block 0 was requested in the prologue
                if (!mbar_wait(&k_empty, (j & 1) ^ 1)) { ok = false; break; }
                mbar_arrive_expect_tx(&k_full, 2 * kTile);
                for (int dc = 0; dc < 2; ++dc)
                    tma_load_3d_hint(smem + kKOff + dc * kTile, &map_qk, &k_full, col_k + dc * 64,
                                s.row0 + j * kKeys, 0, p.dead_policy);
            }
            if (!ok) atomicExch(p.status, kStatusAttnTimeout);
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------ TMA producer: V blocks (own thread,
        // so that a K block never queues behind the release of the V slot)
        if (lane == 0 && nb_max > 0) {
            prefetch_tensormap(&map_v);
            const int col_v = 2 * p.H + head * kD;
            bool ok = true;
            for (int j = 0; j < nb_max && ok; ++j) {
                if (!mbar_wait(&v_empty, (j & 1) ^ 1)) { ok = false; break; }
                mbar_arrive_expect_tx(&v_full, p.v_planes * 2 * kTile);
                for (int dh = 0; dh < 2; ++dh)
                    tma_load_3d_hint(smem + kVOff + dh * 2 * kTile, &map_v, &v_full, col_v + dh * 64,
                                s.row0 + j * kKeys, 0, p.dead_policy);
            }
            if (!ok) atomicExch(p.status, kStatusAttnTimeout);
        }
    } else if (warp == 1 || warp == 2) {
        // ------------------------------------------------------------ MMA issuer of lane L
        const int L = warp - 1;
        if (lane == 0 && nb_max > 0) {
            constexpr uint32_t idesc_s = make_idesc_f16(128, kKeys, false);
            constexpr uint32_t idesc_o = make_idesc_f16(128, kD, true);
            const uint32_t q_addr = smem_u32(smem + kQOff + L * 2 * kTile);
            const uint32_t p_addr = smem_u32(smem + kPOff + L * 2 * kTile);
            const uint32_t k_addr = smem_u32(smem + kKOff), v_addr = smem_u32(smem + kVOff);
            const uint32_t s_tmem = tmem_base + L * 256, o_tmem = s_tmem + 128;
            const int n = L ? nb_1 : nb_0;
            bool ok = true;
            long long tw[5] = {0, 0, 0, 0, 0};
            const long long t_begin = clock64();
            auto timed = [&](uint64_t* bar, uint32_t parity, int slot) {
                const long long t0 = clock64();
                const bool got = mbar_wait(bar, parity);
                tw[slot] += clock64() - t0;
                return got;
            };
            auto issue_s = [&](int j) -> bool {
                if (j > 0 && !timed(&s_empty[L], (j - 1) & 1, 1)) return false;   // scores j-1 are in registers
                if (!timed(&k_full, j & 1, 2)) return false;
                tcgen05_fence_after();
#pragma unroll
                for (int ks = 0; ks < kD / 16; ++ks) {
                    const uint32_t off = (ks >> 2) * kTile + (ks & 3) * 32;
                    umma_f16(s_tmem, smem_desc_kmajor_sw128(q_addr + off), smem_desc_kmajor_sw128(k_addr + off),
                             idesc_s, ks > 0);
                }
                umma_commit(&k_empty);
                umma_commit(&s_full[L]);
                return true;
            };
            if (n > 0) {
                ok = timed(&q_full, 0, 0);
                if (ok) ok = issue_s(0);
                for (int j = 0; j < n && ok; ++j) {
                    if (j + 1 < n && !issue_s(j + 1)) { ok = false; break; }
                    if (!timed(&p_full[L], j & 1, 3) || !timed(&v_full, j & 1, 4)) { ok = false; break; }
                    tcgen05_fence_after();
#pragma unroll
                    for (int ks = 0; ks < kKeys / 16; ++ks) {
                        const uint32_t pa = p_addr + (ks >> 2) * kTile + (ks & 3) * 32;   // 16 keys of the swizzle row
                        const uint32_t vb = v_addr + ks * 16 * 128;                       // 16 key rows
                        const uint64_t dp = smem_desc_kmajor_sw128(pa);
                        umma_f16(o_tmem, dp, smem_desc_mnmajor_sw128(vb, 2 * kTile), idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
                        if (p.v_planes == 2)
                            umma_f16(o_tmem, dp, smem_desc_mnmajor_sw128(vb + kTile, 2 * kTile), idesc_o, 1);
                    }
                    umma_commit(&v_empty);
                    umma_commit(&o_done[L]);
                }
            }
            // a lane with fewer key blocks than its partner (causal mask, idle lane) still owes the
            // shared K / V slots one release per block: one phase at a time, so that two of its
            // arrivals can never complete a phase the partner has not reached
            for (int j = n; j < nb_max && ok; ++j) {
                if (j > 0 && !mbar_wait(&k_empty, (j - 1) & 1)) { ok = false; break; }
                mbar_arrive(&k_empty);
                if (j > 0 && !mbar_wait(&v_empty, (j - 1) & 1)) { ok = false; break; }
                mbar_arrive(&v_empty);
            }
            if (!ok) atomicExch(p.status, kStatusAttnTimeout);
            if (p.trace && L == 0 && n > 0) {
                for (int i = 0; i < 5; ++i) atomicAdd(p.trace + i, (unsigned long long)tw[i]);
                atomicAdd(p.trace + 5, (unsigned long long)(clock64() - t_begin));
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        // ------------------------------------------------------------ softmax + epilogue of lane L
        const int L = (warp - 4) >> 2, quad = warp & 3;
        const int q0 = L ? q0_1 : q0_0, n = L ? nb_1 : nb_0, nk = L ? nkeys_1 : nkeys_0;
        const bool active = L ? active_1 : active_0;
        const int r = quad * 32 + lane, t = q0 + r;
        const int64_t out_row0 = (int64_t)(s.row0 + q0);
        if (active && n == 0) {
            // every key masked: zeros
            for (int i = (warp & 3) * 32 + lane; i < 128 * kD / 8; i += 128) {
                const int row = i / (kD / 8), u = i - row * (kD / 8);
                __half* dst = p.out + (out_row0 + row) * p.H + head * kD + u * 8;
                *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(dst + p.out_plane_stride) = make_uint4(0, 0, 0, 0);
            }
        } else if (n > 0) {
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + L * 256;
            const uint32_t o_row = t_row + 128;
            const uint32_t p_tile = smem_u32(smem + kPOff + L * 2 * kTile);
            const uint32_t q_tile = smem_u32(smem + kQOff + L * 2 * kTile);
            const uint32_t row_off = (uint32_t)r * 128, sw = (uint32_t)(r & 7);
            const float c = p.scale_log2e;
            float m_ref = 0.f, sum = 0.f;
            bool ok = true;
            long long t_s = 0, t_o = 0, t_first = 0;
            const long long t_loop = clock64();
#pragma unroll 1
            for (int j = 0; j < n && ok; ++j) {
                {
                    const long long t0 = clock64();
                    ok = mbar_wait(&s_full[L], j & 1);
                    t_s += clock64() - t0;
                    if (j == 0) t_first = clock64() - t_kernel;
                    if (!ok) break;
                }
                tcgen05_fence_after();
                uint32_t sc[4][32];
                tmem_ld_32x32(t_row, sc[0]);
                tmem_ld_32x32(t_row + 32, sc[1]);
                tmem_ld_32x32(t_row + 64, sc[2]);
                tmem_ld_32x32(t_row + 96, sc[3]);
                tmem_wait_ld();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[L]);   // the next block's scores may overwrite the columns
                const int key0 = j * kKeys;
                const bool full = key0 + kKeys <= nk && (!p.causal || key0 + kKeys - 1 <= q0);
                if (!full) {
#pragma unroll
                    for (int i = 0; i < 128; ++i) {
                        const int key = key0 + i;
                        const bool allowed = key < nk && (!p.causal || key <= t);
                        if (!allowed) sc[i >> 5][i & 31] = __float_as_uint(-FLT_MAX);
                    }
                }
                float bm4[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};   // four independent chains
#pragma unroll
                for (int i = 0; i < 128; ++i) bm4[i & 3] = fmaxf(bm4[i & 3], __uint_as_float(sc[i >> 5][i & 31]));
                const float bm = fmaxf(fmaxf(bm4[0], bm4[1]), fmaxf(bm4[2], bm4[3]));
                // lazy running max: move it only when this block exceeds it by more than 2^8
                float factor = 1.f;
                if (j == 0) {
                    m_ref = bm;
                } else if ((bm - m_ref) * c > kRescaleLog2) {
                    factor = ex2((m_ref - bm) * c);
                    m_ref = bm;
                }
                const bool rescale = __any_sync(0xffffffffu, factor != 1.f);
                bool waited = false;
                if (rescale) {
                    // O may only be touched once P.V of block j-1 has retired
                    if (!mbar_wait(&o_done[L], (j - 1) & 1)) { ok = false; break; }
                    waited = true;
                    tcgen05_fence_after();
                    sum *= factor;
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t o[32];
                        tmem_ld_32x32(o_row + ch * 32, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                        tmem_st_32x32(o_row + ch * 32, o);
                    }
                    tmem_wait_st();
                }
                const float mc = m_ref * c;
                // p = exp2(s c - m c) (masked scores are -FLT_MAX: exp2 flushes to 0), rounded to
                // fp16 for the MMA.  The row sum is taken over the ROUNDED numerators, so the weights
                // the P.V MMA applies sum to one exactly: with few unmasked keys (the first rows of a
                // causal sequence) the rounding of a dominant numerator would otherwise scale the
                // whole output row by 1 + 2^-12 (measured: the causal golden misses 1e-4 with the
                // fp32 sum).  HADD2.F32 unpacks on the fp16 pipe, not on the MUFU pipe.
                uint32_t h[64];
                float sum2 = 0.f;
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const float e0 = ex2(fmaf(__uint_as_float(sc[(2 * i) >> 5][(2 * i) & 31]), c, -mc));
                    const float e1 = ex2(fmaf(__uint_as_float(sc[(2 * i + 1) >> 5][(2 * i + 1) & 31]), c, -mc));
                    h[i] = pack_f16x2(e0, e1);
                    const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
                    sum += back.x;
                    sum2 += back.y;
                }
                sum += sum2;
                // the P tile is free once P.V of block j-1 has retired (the exps above overlap it)
                if (j > 0 && !waited) {
                    const long long t0 = clock64();
                    ok = mbar_wait(&o_done[L], (j - 1) & 1);
                    t_o += clock64() - t0;
                    if (!ok) break;
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    // 8 keys = one 16-byte unit of the 128B-swizzled row; 64 keys per tile
                    const uint32_t addr = p_tile + (uint32_t)(u >> 3) * kTile + row_off + ((((uint32_t)u & 7) ^ sw) << 4);
                    st_shared_v4(addr, h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]);
                }
                fence_proxy_async_smem();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[L]);
            }
            {
                const long long t0 = clock64();
                if (ok && !mbar_wait(&o_done[L], (n - 1) & 1)) ok = false;
                t_o += clock64() - t0;
            }
            const long long t_epi = clock64();
            tcgen05_fence_after();
            if (ok) {
                // O / sum -> split planes, staged as [128 rows][64 cols] 128B-swizzled tiles in the
                // lane's dead Q tile (hi plane) and P tile (lo plane), stored with TMA
                const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t o[32];
                    tmem_ld_32x32(o_row + ch * 32, o);
                    tmem_wait_ld();
                    uint32_t h[16], l[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        split2_f16(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv, h[i], l[i]);
                    const uint32_t tile_off = (uint32_t)(ch >> 1) * kTile + row_off;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t unit = ((uint32_t)((ch & 1) * 4 + u) ^ sw) << 4;
                        st_shared_v4(q_tile + tile_off + unit, h[4 * u], h[4 * u + 1], h[4 * u + 2], h[4 * u + 3]);
                        st_shared_v4(p_tile + tile_off + unit, l[4 * u], l[4 * u + 1], l[4 * u + 2], l[4 * u + 3]);
                    }
                }
                fence_proxy_async_smem();
            }
            named_bar_sync(1 + L, 128);
            if (ok && quad == 0 && lane == 0) {
#pragma unroll
                for (int dc = 0; dc < 2; ++dc) {
                    tma_store_3d(&map_out, smem + kQOff + (L * 2 + dc) * kTile, head * kD + dc * 64, (int)out_row0, 0);
                    tma_store_3d(&map_out, smem + kPOff + (L * 2 + dc) * kTile, head * kD + dc * 64, (int)out_row0, 1);
                }
                bulk_commit_group();
                bulk_wait_read_all();   // shared memory must outlive the reads; the writes complete with the grid
            }
            if (!ok) atomicExch(p.status, kStatusAttnTimeout);
            if (p.trace && L == 0 && quad == 0 && lane == 0) {
                atomicAdd(p.trace + 8, (unsigned long long)t_s);
                atomicAdd(p.trace + 9, (unsigned long long)t_o);
                atomicAdd(p.trace + 10, (unsigned long long)(t_epi - t_loop));
                atomicAdd(p.trace + 11, (unsigned long long)(clock64() - t_epi));
                atomicAdd(p.trace + 12, (unsigned long long)t_first);
                atomicAdd(p.trace + 13, (unsigned long long)(clock64() - t_kernel));
                atomicAdd(p.trace + 14, 1ull);
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 3) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace

bool attention_dual_supported(int D, int max_pitch, int qk_planes, int p_planes) {
    return D == kD && qk_planes == 1 && p_planes == 1 && max_pitch % 128 == 0;
}

int launch_attention_dual(ppgs_engine* e, const __half* qkv, __half* out, int rows, int H, int heads,
                          int max_pitch, int nseq, const SeqInfo* seqs_dev, int causal, int planes,
                          cudaStream_t stream, int q_first_tile, int q_tiles) {
    const bool per_seq = q_first_tile < 0;
    CUtensorMap map_qk, map_v, map_out;
    PPGS_CHECK(make_store_map(&map_out, out, H, rows, (uint64_t)rows * H));
    PPGS_CHECK(make_plane_map(&map_qk, qkv, false, 3 * H, rows, 1, 2, 3 * H, 0, (uint64_t)rows * 3 * H, 128, 1));
    PPGS_CHECK(make_plane_map(&map_v, qkv, false, 3 * H, rows, 1, 2, 3 * H, 0, (uint64_t)rows * 3 * H, 128,
                              planes));
    static PerDeviceOnce attr;
    if (attr.first(e->device)) {
        PPGS_CUDA(cudaFuncSetAttribute(attention_dual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kDualSmem));
    }
    DualParams p;
    p.seqs = seqs_dev;
    p.H = H;
    p.causal = causal;
    p.reverse = e->attn_reverse;
    p.dead_policy = (e->l2_hints && !per_seq) ? kL2EvictFirst : kL2EvictNormal;   // the streaming decoder re-reads its K / V caches
    p.v_planes = planes;
    p.scale_log2e = 1.4426950408889634f / sqrtf((float)kD);
    p.out = out;
    p.out_plane_stride = (int64_t)rows * H;
    p.q_first_tile = per_seq ? 0 : q_first_tile;
    p.q_first_per_seq = per_seq ? 1 : 0;
    p.q_tiles = q_tiles;
    p.status = e->status_dev;
    p.trace = e->trace_dev ? e->trace_dev + 80 : nullptr;
    const int tiles = q_tiles > 0 ? q_tiles : max_pitch / 128;
    dim3 grid((tiles + 1) / 2, heads, (unsigned)nseq);
    {
        LaunchScope scope(e, "tc_attention", stream);
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attrs[1];
        cfg.gridDim = grid;
        cfg.blockDim = dim3(kDualThreads);
        cfg.dynamicSmemBytes = kDualSmem;
        cfg.stream = stream;
        if (pdl_enabled()) {
            attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attrs[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attrs;
            cfg.numAttrs = 1;
        }
        PPGS_CUDA(cudaLaunchKernelEx(&cfg, attention_dual_kernel, map_qk, map_v, map_out, p));
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

}  // namespace ppgs
