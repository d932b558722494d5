// Attention over split-fp16 planes (attention_tc.cu).
#pragma once
#include "common.cuh"

namespace ppgs {

// qkv: planes [2][rows][3H] (Q | K | V, heads contiguous inside each), out: planes
// [2][rows][H].  Key-padding mask from SeqInfo::valid_len, optional causal mask
// (ppgs/model/transformer.py:65-80), fp32 softmax.
int launch_attention_tc(ppgs_engine* e, const __half* qkv, __half* out, int rows,
                        const ForwardPlan& plan, const SeqInfo* seqs_dev, int planes,
                        cudaStream_t stream);

// Same for any (hidden, heads): tcgen05 kernel when head_dim is 64 / 128 / 256 and the
// pitch is <= 512 rows, CUDA-core kernel otherwise.
// `q_first_tile` / `q_tiles` restrict the query tiles of every sequence (streaming decoder: the
// keys of earlier tiles are the cache); 0 / 0 = all; q_first_tile = -1: the first query tile of
// every sequence comes from its SeqInfo::src_start.
int launch_attention_any(ppgs_engine* e, const __half* qkv, __half* out, int rows, int H, int heads,
                         int max_pitch, int nseq, const SeqInfo* seqs_dev, int causal, int planes,
                         cudaStream_t stream, int q_first_tile = 0, int q_tiles = 0,
                         int qk_planes = 2, int p_planes = 2);
// `qk_planes` / `p_planes` = 1: Q, K / the softmax numerators P enter their MMAs as one fp16
// plane (S in one pass instead of three, P.V in two); ignored when `planes` is 1.

// CUDA-core attention over split planes for any (hidden, heads, head_dim in {64, 128, 256}):
// used by the wav2vec2 encoder (12 heads x 64).
int launch_attention_planes(ppgs_engine* e, int head_dim, const __half* qkv, __half* out, int rows,
                            int H, int heads, int max_pitch, int nseq, const SeqInfo* seqs_dev,
                            int causal, int planes, cudaStream_t stream);

// Two query tiles per CTA sharing every K / V block, single-plane Q / K / P (attention_dual_tc.cu):
// head_dim 128 only.  Same arguments as launch_attention_any.
bool attention_dual_supported(int D, int max_pitch, int qk_planes, int p_planes);
int launch_attention_dual(ppgs_engine* e, const __half* qkv, __half* out, int rows, int H, int heads,
                          int max_pitch, int nseq, const SeqInfo* seqs_dev, int causal, int planes,
                          cudaStream_t stream, int q_first_tile, int q_tiles);

}  // namespace ppgs
