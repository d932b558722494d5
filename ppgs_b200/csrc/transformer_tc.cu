// Transformer forward on the tensor cores (PPGS_PRECISION_F16X2 / _F16):
// ppgs/model/transformer.py:45-81 + softmax (ppgs/core.py:593-594) as a short
// sequence of fused tcgen05 kernels over the folded-sequence row layout.
//
//   fold          (B,C,T) fp16 features -> time-major fp16 rows (chunk fold, replicate pad)
//   conv_in       5-tap GEMM  + bias + length mask + positional encoding   -> x
//   per layer     QKV GEMM + bias -> split planes
//                 attention (length / causal mask, fp32 softmax)
//                 out-proj GEMM + bias + residual + LayerNorm              -> x
//                 linear1 GEMM + bias + ReLU -> split planes
//                 linear2 GEMM + bias + residual + LayerNorm               -> x
//   conv_out      5-tap GEMM + bias + mask + channel softmax + un-chunk    -> (B,40,T)
//
// Every activation lives in HBM as split-fp16 planes [2][rows][K] (hi + lo carries
// 22 significand bits), including the residual stream.
#include <stdlib.h>

#include <algorithm>

#include "attention_tc.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace ppgs {

using namespace tc;

int build_weight_map(TcWeight& w) {
    const uint64_t plane = (uint64_t)w.taps * w.N * w.C;
    for (int planes = 1; planes <= 2; ++planes) {
        TcWeight::Maps& m = w.maps[planes - 1];
        const struct { CUtensorMap* map; uint32_t rows; } boxes[4] = {
            {&m.bn256, 256}, {&m.bn128, 128}, {&m.bn64, 64}, {&m.bn32, 32}};
        for (const auto& box : boxes)
            PPGS_CHECK(make_plane_map(box.map, w.planes, true, w.C, w.N, w.taps, 2, w.C,
                                      (uint64_t)w.N * w.C, plane, box.rows, planes));
    }
    return PPGS_OK;
}

int build_weight_maps(ppgs_engine* e) {
    if (e->tc_maps_ready) return PPGS_OK;
    PPGS_CHECK(build_weight_map(e->tc_conv_in));
    PPGS_CHECK(build_weight_map(e->tc_conv_out));
    for (TcLayer& l : e->tc_layers) {
        PPGS_CHECK(build_weight_map(l.in_w));
        PPGS_CHECK(build_weight_map(l.out_w));
        PPGS_CHECK(build_weight_map(l.l1_w));
        PPGS_CHECK(build_weight_map(l.l2_w));
    }
    PPGS_CHECK(ensure_status_word(e));
    e->tc_maps_ready = true;
    return PPGS_OK;
}

int ensure_status_word(ppgs_engine* e) {
    if (!e->status_dev) {
        PPGS_CUDA(cudaMalloc(&e->status_dev, sizeof(int)));
        PPGS_CUDA(cudaMemset(e->status_dev, 0, sizeof(int)));
    }
    return PPGS_OK;
}

int check_status(ppgs_engine* e, cudaStream_t stream) {
    if (!e->status_dev) return PPGS_OK;
    int status = 0;
    PPGS_CUDA(cudaMemcpyAsync(&status, e->status_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
    PPGS_CUDA(cudaStreamSynchronize(stream));
    if (status != 0) {
        cudaMemsetAsync(e->status_dev, 0, sizeof(int), stream);
        set_error("tensor-core pipeline timed out waiting on a barrier (status %d)", status);
        return PPGS_E_CUDA;
    }
    return PPGS_OK;
}

// x <- LayerNorm(x + y) over split planes, one warp per row; rows outside the sequence
// tensor -> 0 (they are the zero halo of the output convolution).  Hidden sizes > 256,
// where the GEMM epilogue cannot see a whole row.
template <int H>
__global__ void __launch_bounds__(256)
residual_layernorm_planes_kernel(__half* __restrict__ x, const __half* __restrict__ y, int64_t plane_stride,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                 const SeqInfo* __restrict__ seqs, const int* __restrict__ tile_seq, int rows) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    constexpr int PER = H / 64;   // half2 per lane
    const SeqInfo s = seqs[tile_seq[row >> 7]];
    const int64_t base = (int64_t)row * H;
    if (row - s.row0 >= s.tensor_len) {
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int64_t at = base + 2 * (lane + 32 * i);
            *reinterpret_cast<uint32_t*>(x + at) = 0u;
            *reinterpret_cast<uint32_t*>(x + plane_stride + at) = 0u;
        }
        return;
    }
    float v[2 * PER];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int64_t at = base + 2 * (lane + 32 * i);
        const float2 xh = __half22float2(*reinterpret_cast<const __half2*>(x + at));
        const float2 xl = __half22float2(*reinterpret_cast<const __half2*>(x + plane_stride + at));
        const float2 yh = __half22float2(*reinterpret_cast<const __half2*>(y + at));
        const float2 yl = __half22float2(*reinterpret_cast<const __half2*>(y + plane_stride + at));
        v[2 * i] = (xh.x + xl.x) + (yh.x + yl.x);
        v[2 * i + 1] = (xh.y + xl.y) + (yh.y + yl.y);
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
        uint32_t hi2, lo2;
        tc::split2_f16((v[2 * i] - mean) * rstd * gamma[c] + beta[c],
                       (v[2 * i + 1] - mean) * rstd * gamma[c + 1] + beta[c + 1], hi2, lo2);
        *reinterpret_cast<uint32_t*>(x + base + c) = hi2;
        *reinterpret_cast<uint32_t*>(x + plane_stride + base + c) = lo2;
    }
}

// (B, C, T) fp16 -> [rows][C] fp16, time-major (transformer.py:54,58 folded)
__global__ void fold_half_kernel(const __half* __restrict__ feats, int C, int T,
                                 const SeqInfo* __restrict__ seqs, const int* __restrict__ tile_seq,
                                 __half* __restrict__ x0) {
    __shared__ __half tile[32][34];
    const int row_base = blockIdx.x * 32, c_base = blockIdx.y * 32;
    const SeqInfo s = seqs[tile_seq[row_base >> 7]];
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int i = ty; i < 32; i += 8) {
        const int c = c_base + i, t = row_base - s.row0 + tx;
        __half v = __float2half_rn(0.f);
        if (c < C && t < s.tensor_len)
            v = feats[((int64_t)s.batch * C + c) * T + max(s.src_start + t, 0)];
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c_base + tx;
        if (c < C) x0[(int64_t)(row_base + i) * C + c] = tile[tx][i];
    }
}

static bool split_out_proj_experiment() {
    // experiment (PPGS_B200_SPLIT_OUT_PROJ_LN=1): the K = 256 out-projection is epilogue-bound
    // with the fused LayerNorm; run it with the plain epilogue + the standalone LayerNorm pass
    static const bool on = [] { const char* v = getenv("PPGS_B200_SPLIT_OUT_PROJ_LN"); return v && atoi(v) != 0; }();
    return on;
}

// Workspace of one forward: offsets of the activation buffers and plan tables.
struct TcWorkspace {
    size_t o_x0, o_xh, o_qkv, o_att, o_ff, o_y, o_seqs, o_tiles, bytes;
};
static TcWorkspace tc_workspace(const ppgs_model_config& c, const ForwardPlan& plan) {
    const size_t rows = (size_t)plan.rows, C = c.input_channels, H = c.hidden_channels, F = c.ffn_channels;
    Carver w;
    TcWorkspace t;
    t.o_x0 = w.take(rows * C * 2);
    t.o_xh = w.take(2 * rows * H * 2);
    t.o_qkv = w.take(2 * rows * 3 * H * 2);
    t.o_att = w.take(2 * rows * H * 2);
    t.o_ff = w.take(2 * rows * F * 2);
    t.o_y = w.take(H == 256 && !split_out_proj_experiment() ? 0 : 2 * rows * H * 2);
    t.o_seqs = w.take(plan.seqs.size() * sizeof(SeqInfo));
    t.o_tiles = w.take((rows / 128) * 4);
    t.bytes = w.off;
    return t;
}

// The input convolution's operand rows ([rows][C] fp16, time-major, zero beyond each chunk
// tensor) and the device plan table, for a producer that writes them itself (the mel kernel)
// and then calls transformer_forward_tc with features == nullptr.
int transformer_tc_input_rows(ppgs_engine* e, const ForwardPlan& plan, cudaStream_t stream, __half** x0,
                              const SeqInfo** seqs_dev) {
    if (!tensor_core_shape(e->cfg)) {
        set_error("tensor-core path does not cover this model shape");
        return PPGS_E_UNSUPPORTED;
    }
    const TcWorkspace w = tc_workspace(e->cfg, plan);
    PPGS_CHECK(ensure_workspace(e, w.bytes));
    char* ws = static_cast<char*>(e->workspace);
    PPGS_CHECK(upload_plan(e, plan, reinterpret_cast<SeqInfo*>(ws + w.o_seqs),
                           reinterpret_cast<int*>(ws + w.o_tiles), stream));
    *x0 = reinterpret_cast<__half*>(ws + w.o_x0);
    *seqs_dev = reinterpret_cast<const SeqInfo*>(ws + w.o_seqs);
    return PPGS_OK;
}

int transformer_forward_tc(ppgs_engine* e, const __half* features, const ForwardPlan& plan,
                           int softmax, float* out, cudaStream_t stream) {
    const ppgs_model_config& c = e->cfg;
    const int C = c.input_channels, H = c.hidden_channels, F = c.ffn_channels;
    const int O = c.output_channels, k = c.kernel_size;
    const int rows = plan.rows, D = H / c.num_heads;
    if (!tensor_core_shape(c)) {
        set_error("tensor-core path supports hidden %% 256 == 0 with head_dim 64 / 128 / 256 (got %d / %d)",
                  H, D);
        return PPGS_E_UNSUPPORTED;
    }
    // hidden 256: one GEMM tile spans a full row and the residual + LayerNorm live in the
    // epilogue; wider models store the projection and normalise in a separate pass
    const bool fused_ln = H == 256;
    const bool split_out_proj = split_out_proj_experiment();
    PPGS_CHECK(build_weight_maps(e));
    const int planes = e->precision == PPGS_PRECISION_F16X2 ? 2 : 1;

    const TcWorkspace w = tc_workspace(c, plan);
    PPGS_CHECK(ensure_workspace(e, w.bytes));
    char* ws = static_cast<char*>(e->workspace);
    __half* x0 = reinterpret_cast<__half*>(ws + w.o_x0);
    __half* xh = reinterpret_cast<__half*>(ws + w.o_xh);
    __half* qkv = reinterpret_cast<__half*>(ws + w.o_qkv);
    __half* att = reinterpret_cast<__half*>(ws + w.o_att);
    __half* ff = reinterpret_cast<__half*>(ws + w.o_ff);
    __half* yh = reinterpret_cast<__half*>(ws + w.o_y);
    SeqInfo* seqs_dev = reinterpret_cast<SeqInfo*>(ws + w.o_seqs);
    int* tile_seq_dev = reinterpret_cast<int*>(ws + w.o_tiles);
    PPGS_CHECK(upload_plan(e, plan, seqs_dev, tile_seq_dev, stream));

    // activation tensor maps (A operands): {K, rows, planes}
    CUtensorMap map_x0, map_x, map_att, map_ff;
    PPGS_CHECK(make_plane_map(&map_x0, x0, false, C, rows, 1, 1, C, 0, (uint64_t)rows * C, 128, 1));
    PPGS_CHECK(make_plane_map(&map_x, xh, false, H, rows, 1, 2, H, 0, (uint64_t)rows * H, 128, planes));
    PPGS_CHECK(make_plane_map(&map_att, att, false, H, rows, 1, 2, H, 0, (uint64_t)rows * H, 128, planes));
    PPGS_CHECK(make_plane_map(&map_ff, ff, false, F, rows, 1, 2, F, 0, (uint64_t)rows * F, 128, planes));
    // residual rows of the fused FFN's LayerNorm epilogue: always both planes of x
    CUtensorMap map_res;
    PPGS_CHECK(make_plane_map(&map_res, xh, false, H, rows, 1, 2, H, 0, (uint64_t)rows * H, 128, 2));
    // output tensor maps (TMA stores of the epilogues)
    CUtensorMap out_x, out_qkv, out_ff, out_y;
    PPGS_CHECK(make_store_map(&out_x, xh, H, rows, (uint64_t)rows * H));
    if (!fused_ln || split_out_proj) PPGS_CHECK(make_store_map(&out_y, yh, H, rows, (uint64_t)rows * H));
    PPGS_CHECK(make_store_map(&out_qkv, qkv, 3 * H, rows, (uint64_t)rows * 3 * H));
    PPGS_CHECK(make_store_map(&out_ff, ff, F, rows, (uint64_t)rows * F));

    if (features) {   // nullptr: the mel kernel already wrote x0 (transformer_tc_input_rows)
        dim3 grid(rows / 32, (C + 31) / 32);
        LaunchScope scope(e, "fold_chunks", stream);
        fold_half_kernel<<<grid, dim3(32, 8), 0, stream>>>(features, C, plan.frames, seqs_dev,
                                                           tile_seq_dev, x0);
    }
    PPGS_CUDA(cudaGetLastError());

    GemmParams base;
    base.m_tiles = rows / 128;
    base.seqs = seqs_dev;
    base.tile_seq = tile_seq_dev;
    base.status = e->status_dev;
    base.eps = c.layer_norm_eps;
    base.b_planes = planes;
    const int pair = (e->gemm_pair && rows % 256 == 0) ? 1 : 0;
    base.pair = pair;
    if (const char* v = getenv("PPGS_B200_DEBUG_FLAGS")) base.debug_flags = atoi(v);   // timing experiments
    auto wmap = [&](TcWeight& w) -> const CUtensorMap& {
        return pair ? w.maps[planes - 1].bn128 : w.maps[planes - 1].bn256;
    };
    // cycle accounting slots: 0 conv_in, 1 qkv, 2 out_proj, 3 ffn1, 4 ffn2, 5 conv_out
    auto trace = [&](int slot) { return e->trace_dev ? e->trace_dev + 8 * slot : nullptr; };

    // Serpentine order: every kernel walks the row tiles in the direction opposite to its
    // predecessor's, so it starts on the rows that kernel wrote last (still in L2).
    int direction = 0;
    auto next_direction = [&]() {
        const int d = e->serpentine ? direction : 0;
        direction ^= 1;
        return d;
    };
    {   // input conv: features are exact fp16 -> one A plane
        GemmParams p = base;
        p.reverse = next_direction();
        p.n_tiles = H / 256; p.taps = k; p.half = k / 2; p.cblocks = (C + 63) / 64; p.a_planes = 1;
        p.N = H; p.scale = e->tc_conv_in.inv_scale; p.bias = e->conv_in_b; p.pe = e->pe;
        p.trace = trace(0);
        PPGS_CHECK(launch_gemm_tc(e, "tc_conv_in", 256, kEpiConvIn, map_x0, wmap(e->tc_conv_in),
                                  &out_x, p, stream));
    }
    for (int layer = 0; layer < c.num_layers; ++layer) {
        const LayerWeights& L = e->layers[layer];
        TcLayer& T = e->tc_layers[layer];
        {
            GemmParams p = base;
            p.n_tiles = 3 * H / 256; p.cblocks = H / 64; p.a_planes = planes;
            p.N = 3 * H; p.scale = T.in_w.inv_scale; p.bias = L.in_b; p.trace = trace(1);
            p.reverse = next_direction();
            // Q / K enter the attention MMAs as their hi planes only: skip the lo-plane stores
            if (planes == 2 && e->attn_qk_planes == 1 && e->attention_impl == 1 &&
                (H / c.num_heads == 128 ? e->attn_dual != 0 || plan.max_pitch <= 512 : plan.max_pitch <= 512) &&
                plan.max_pitch % 128 == 0)
                p.hi_only_cols = 2 * H, p.hi_only_passes = e->qk_gemm_passes;
            PPGS_CHECK(launch_gemm_tc(e, "tc_qkv", 256, kEpiPlanes, map_x, wmap(T.in_w), &out_qkv,
                                      p, stream));
        }
        e->attn_reverse = next_direction();
        PPGS_CHECK(launch_attention_tc(e, qkv, att, rows, plan, seqs_dev, planes, stream));
        // x <- LayerNorm(x + projection): epilogue of the GEMM (hidden 256) or its own pass
        auto project_ln = [&](const char* name, const CUtensorMap& map_a, TcWeight& wt, int cblocks,
                              const float* bias, const float* gamma, const float* beta, int slot,
                              int slot_ln) -> int {
            GemmParams p = base;
            p.cblocks = cblocks; p.a_planes = planes;
            p.N = H; p.scale = wt.inv_scale; p.bias = bias; p.trace = trace(slot);
            p.reverse = next_direction();
            if (fused_ln && pair && e->proj_ln && !(split_out_proj && slot == 2)) {
                // CTA-pair projection kernel with the register-resident LayerNorm epilogue
                FfnParams f;
                f.m_tiles = rows / 128; f.num_chunks = cblocks; f.planes = planes;
                f.scale1 = wt.inv_scale; f.scale2 = wt.inv_scale;
                f.bias1 = bias; f.bias2 = bias; f.gamma = gamma; f.beta = beta;
                f.eps = c.layer_norm_eps; f.seqs = seqs_dev; f.tile_seq = tile_seq_dev;
                f.status = e->status_dev; f.trace = nullptr; f.reverse = p.reverse;
                if (e->l2_hints) f.dead_policy = 0x12F0000000000000ull;   // attention output + old x: dead after this kernel
                return launch_proj_ln(e, name, map_a, wt.maps[planes - 1].bn128, out_x, map_res, f, stream);
            }
            if (fused_ln && !(split_out_proj && slot == 2)) {
                p.n_tiles = 1; p.trace_ln = trace(slot_ln);
                p.residual = xh; p.res_ld = H; p.res_plane_stride = (int64_t)rows * H;
                p.gamma = gamma; p.beta = beta;
                return launch_gemm_tc(e, name, 256, kEpiResLN, map_a, wmap(wt), &out_x, p, stream);
            }
            p.n_tiles = H / 256;
            PPGS_CHECK(launch_gemm_tc(e, name, 256, kEpiPlanes, map_a, wmap(wt), &out_y, p, stream));
            LaunchScope scope(e, "residual_layernorm_planes", stream);
            if (H == 256)
                residual_layernorm_planes_kernel<256><<<(rows + 7) / 8, 256, 0, stream>>>(
                    xh, yh, (int64_t)rows * H, gamma, beta, c.layer_norm_eps, seqs_dev, tile_seq_dev, rows);
            else if (H == 512)
                residual_layernorm_planes_kernel<512><<<(rows + 7) / 8, 256, 0, stream>>>(
                    xh, yh, (int64_t)rows * H, gamma, beta, c.layer_norm_eps, seqs_dev, tile_seq_dev, rows);
            else if (H == 768)
                residual_layernorm_planes_kernel<768><<<(rows + 7) / 8, 256, 0, stream>>>(
                    xh, yh, (int64_t)rows * H, gamma, beta, c.layer_norm_eps, seqs_dev, tile_seq_dev, rows);
            else
                residual_layernorm_planes_kernel<1024><<<(rows + 7) / 8, 256, 0, stream>>>(
                    xh, yh, (int64_t)rows * H, gamma, beta, c.layer_norm_eps, seqs_dev, tile_seq_dev, rows);
            PPGS_CUDA(cudaGetLastError());
            return PPGS_OK;
        };
        PPGS_CHECK(project_ln(fused_ln ? "tc_out_proj_ln" : "tc_out_proj", map_att, T.out_w, H / 64, L.out_b,
                              L.n1_w, L.n1_b, 2, 6));
        // The fused kernel walks the 16 hidden chunks of a row-tile pair on ONE CTA pair; with
        // fewer pair tiles than half of the CTA-pair slots the two-GEMM form is shorter (its
        // linear1 spreads the same tile over 8 CTA pairs): 37 pair tiles = 9 472 rows on a B200.
        const bool enough_tiles = rows / 256 >= e->sm_count / 4 || e->fused_ffn == 2;
        if (pair && e->fused_ffn && enough_tiles && fused_ln && F % 128 == 0) {
            FfnParams f;
            f.m_tiles = rows / 128; f.num_chunks = F / 128; f.planes = planes;
            f.scale1 = T.l1_w.inv_scale; f.scale2 = T.l2_w.inv_scale;
            f.bias1 = L.l1_b; f.bias2 = L.l2_b; f.gamma = L.n2_w; f.beta = L.n2_b;
            f.eps = c.layer_norm_eps; f.seqs = seqs_dev; f.tile_seq = tile_seq_dev;
            f.status = e->status_dev;
            f.trace = e->trace_dev ? e->trace_dev + 64 : nullptr;   // counters 64..79
            f.reverse = next_direction();
            if (e->l2_hints) f.dead_policy = 0x12F0000000000000ull;   // residual rows are overwritten by the outputs
            PPGS_CHECK(launch_ffn_fused(e, map_x, T.l1_w.maps[planes - 1].bn64, T.l2_w.maps[planes - 1].bn128, out_x,
                                        map_res, f, stream));
        } else {
            {
                GemmParams p = base;
                p.n_tiles = F / 256; p.cblocks = H / 64; p.a_planes = planes;
                p.N = F; p.scale = T.l1_w.inv_scale; p.bias = L.l1_b; p.relu = 1; p.trace = trace(3);
                p.reverse = next_direction();
                PPGS_CHECK(launch_gemm_tc(e, "tc_ffn1", 256, kEpiPlanes, map_x, wmap(T.l1_w), &out_ff,
                                          p, stream));
            }
            PPGS_CHECK(project_ln(fused_ln ? "tc_ffn2_ln" : "tc_ffn2", map_ff, T.l2_w, F / 64, L.l2_b, L.n2_w,
                                  L.n2_b, 4, 7));
        }
    }
    {
        GemmParams p = base;
        p.n_tiles = 1; p.taps = k; p.half = k / 2; p.cblocks = H / 64; p.a_planes = planes;
        p.N = O; p.O = O; p.scale = e->tc_conv_out.inv_scale; p.bias = e->conv_out_b;
        p.ppg = out; p.T = plan.frames; p.softmax = softmax; p.pair = 0;
        p.reverse = next_direction();
        PPGS_CHECK(launch_gemm_tc(e, "tc_conv_out_softmax", 64, kEpiConvOut, map_x,
                                  e->tc_conv_out.maps[planes - 1].bn64, nullptr, p, stream));
    }
    e->attn_reverse = 0;
    return PPGS_OK;
}

}  // namespace ppgs
