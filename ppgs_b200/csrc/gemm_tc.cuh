// Host-visible interface of the tcgen05 GEMM (gemm_tc.cu).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace ppgs {
namespace tc {

constexpr int kBM = 128;   // rows per tile = TMEM lanes
constexpr int kBK = 64;    // fp16 elements per k-block = one 128-byte swizzle row

enum GemmEpilogue : int {
    kEpiF32 = 0,      // fp32 [M][N] = act(acc*scale + bias): validation, grouped positional conv
    kEpiPlanes = 1,   // bias (+ReLU) -> split-fp16 planes [2][M][N]
    kEpiConvIn = 2,   // input conv: bias, length mask, + positional encoding -> x planes
    kEpiResLN = 3,    // bias + residual + LayerNorm -> x planes (in place)
    kEpiConvOut = 4,  // output conv: bias, mask, channel softmax, un-chunk -> (B, O, T)
};

struct GemmParams {
    int m_tiles = 0, n_tiles = 0;
    int reverse = 0;             // walk the row tiles from the last to the first (serpentine order across the kernels of a forward: start where the previous kernel's data is still in L2)
    int taps = 1, half = 0;      // K loop = taps x cblocks k-blocks; A rows shift by tap - half
    int row_mul = 1;             // A row of output row m = m * row_mul + tap - half (strided conv)
    int cblocks = 0;
    int cb0 = 0;                 // first k-block: the GEMM covers K columns [64 cb0, 64 (cb0 + cblocks)) of A and W
    // grouped GEMM (block-diagonal weights): n tile g reads A rows shifted by g * a_group_rows
    // and owns output columns [g * group_cols, (g + 1) * group_cols) (kEpiF32 only)
    int a_group_rows = 0, group_cols = 0;
    // row-tile window (streaming decoder): only tiles [win_first, win_first + win_size) of every
    // sequence of `win_stride` tiles are computed (m_tiles = sequences * win_size; win_size even
    // for the CTA-pair kernel); 0 = all rows
    int win_size = 0, win_stride = 0, win_first = 0;
    int win_per_seq = 0;         // 1: the first tile of sequence s is seqs[s].src_start instead of win_first
    int a_planes = 2, b_planes = 2;
    int pair = 0;                // 1: CTA-pair kernel (cta_group::2); W map must have box rows BN/2
    int N = 0;                   // real output columns
    const float* scale = nullptr;   // device scalar: 1 / (power-of-two weight scale)
    const float* bias = nullptr;
    float* out_f32 = nullptr;       // kEpiF32 only
    int64_t ld_f32 = 0;
    // plane outputs leave through `map_out` (TMA store); the residual stream is
    // read back from its planes [2][rows][res_ld]
    const __half* residual = nullptr;
    int64_t res_ld = 0, res_plane_stride = 0;
    const float* gamma = nullptr;
    const float* beta = nullptr;
    float eps = 1e-5f;
    const float* pe = nullptr;
    const SeqInfo* seqs = nullptr;
    const int* tile_seq = nullptr;
    int relu = 0;                // kEpiPlanes activation: 0 none, 1 ReLU, 2 exact GELU
    // CTA-pair kEpiPlanes only: the hi.hi products accumulate in one TMEM buffer and the two
    // correction products in the other (no tile overlap).  tcgen05 accumulation truncates, so the
    // error of an accumulator grows with its number of MMA steps; the large sum then sees one
    // third of them.  Used where K is long and the tolerance is a flip of an fp16 feature.
    int split_acc = 0;
    int hi_only_passes = 3;      // MMA passes for the N blocks that lie below hi_only_cols: 3 (hi.hi + both corrections), 2 (hi.hi + a_lo.b_hi) or 1
    int hi_only_cols = 0;        // kEpiPlanes: output columns [0, hi_only_cols) keep only their hi plane (multiple of 64)
    float* ppg = nullptr;   // kEpiConvOut
    int T = 0, O = 0, softmax = 1;
    int* status = nullptr;
    // optional cycle accounting (validation builds of the pipeline, see debug_abi.h):
    // [0] MMA wait tmem_empty [1] MMA wait full [2] MMA total [3] epilogue wait tmem_full
    // [4] epilogue total [5] producer wait empty [6] producer total [7] CTAs
    unsigned long long* trace = nullptr;
    unsigned long long* trace_ln = nullptr;   // kEpiResLN phases: pass1, exchange, pass2, exchange, pass3
    int debug_flags = 0;   // timing experiments only: 1 = skip plane stores, 2 = skip parameter loads
};

// Tensor map over split-fp16 planes: logical dims {inner, rows, groups, planes}
// (groups = conv taps for weights), box {64, box_rows, 1, box_planes}, 128-byte
// swizzle, zero fill out of bounds.  Activation maps are rank 3 {inner, rows,
// planes} (`rank4` false: groups must be 1); weight maps are always rank 4.
int make_plane_map(CUtensorMap* map, const __half* base, bool rank4, uint64_t inner,
                   uint64_t rows, uint64_t groups, uint64_t planes, uint64_t row_stride_elems,
                   uint64_t group_stride_elems, uint64_t plane_stride_elems, uint32_t box_rows,
                   uint32_t box_planes, uint32_t row_step = 1);
// `row_step` > 1: the box takes every row_step-th row (TMA element stride) — the A operand
// of a strided convolution over time-major activations (rank-3 maps only).

// BN in {256, 128, 64}.  Launches a persistent grid of min(tiles, SMs) CTAs.
// PPGS_B200_PDL=1: GEMM and attention kernels are launched with programmatic stream serialisation
// (their prologues overlap the previous kernel's tail; see pdl_wait in tc_common.cuh)
bool pdl_enabled();

// Store map over output planes [2][rows][inner]: box {64, 128, 1 plane}, 128-byte
// swizzle (the epilogue's staging layout).
int make_store_map(CUtensorMap* map, __half* base, uint64_t inner, uint64_t rows,
                   uint64_t plane_stride_elems);

// `map_out` may be null for epilogues without plane output (kEpiF32, kEpiConvOut).
int launch_gemm_tc(ppgs_engine* e, const char* name, int bn, int epilogue,
                   const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap* map_out,
                   const GemmParams& p, cudaStream_t stream);

// Fused linear1 + ReLU + linear2 + residual + LayerNorm (ffn_tc.cu), CTA pairs.
struct FfnParams {
    int m_tiles, num_chunks, planes;
    int reverse = 0;             // as GemmParams::reverse
    unsigned long long dead_policy = 0x1000000000000000ull;   // L2 policy of loads whose source is dead after the kernel (residual rows; the projection's A operand): tc_common.cuh kL2Evict*
    // row-tile window (streaming decoder, same meaning as GemmParams::win_*): m_tiles = sequences x
    // win_size, the first tile of sequence s is seqs[s].src_start; 0 = all rows
    int win_size = 0, win_stride = 0;
    const float* scale1;
    const float* scale2;
    const float* bias1;
    const float* bias2;
    const float* gamma;
    const float* beta;
    float eps;
    const SeqInfo* seqs;
    const int* tile_seq;
    int* status;
    // cycle accounting (PPGS_B200_TRACE=1): MMA waits [0] x_full [1] hacc_empty [2] w1_full
    // [3] h1_full [4] w2_full [5] y_empty [6] total; epilogue [8] hacc_full [9] h1_empty
    // [10] y_full [11] LayerNorm [12] total [13] CTAs
    unsigned long long* trace;
};
// map_x: activation planes {256, rows, 2}, box {64, 128, planes}; map_w1: linear1 planes,
// box rows 64; map_w2: linear2 planes, box rows 128; map_out: store map of the x planes;
// map_res: the x planes again with both planes in the box (residual rows of the epilogue).
// num_chunks = ffn_channels / 128.
int launch_ffn_fused(ppgs_engine* e, const CUtensorMap& map_x, const CUtensorMap& map_w1,
                     const CUtensorMap& map_w2, const CUtensorMap& map_out, const CUtensorMap& map_res,
                     const FfnParams& p, cudaStream_t stream);

// x <- LayerNorm(x + A . W^T + bias2) for hidden 256 (ffn_tc.cu, proj_ln_kernel): CTA pairs, accumulator
// released after one read, residual by TMA.  Uses FfnParams: num_chunks = K / 64, scale2 / bias2 /
// gamma / beta / eps / seqs / tile_seq / status / planes / m_tiles.  map_a: A planes, box {64, 128,
// planes}; map_w: weight planes, box rows 128.
int launch_proj_ln(ppgs_engine* e, const char* name, const CUtensorMap& map_a, const CUtensorMap& map_w,
                   const CUtensorMap& map_out, const CUtensorMap& map_res, const FfnParams& p,
                   cudaStream_t stream);

}  // namespace tc
}  // namespace ppgs
