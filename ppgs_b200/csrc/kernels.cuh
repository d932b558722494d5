// Internal launch helpers shared by the transformer paths.
#pragma once
#include "common.cuh"

namespace ppgs {

// Bump allocator over the engine workspace (256-byte aligned slices).
struct Carver {
    size_t off = 0;
    size_t take(size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~size_t(255);
        return o;
    }
};

// Copies the plan tables (SeqInfo per sequence, sequence index per 128-row tile)
// to the device through the engine's pinned staging buffer.
int upload_plan(ppgs_engine* e, const ForwardPlan& plan, SeqInfo* seqs_dev, int* tile_seq_dev,
                cudaStream_t stream);

int launch_fold(ppgs_engine* e, const __half* feats, const ForwardPlan& plan,
                const SeqInfo* seqs_dev, const int* tile_seq_dev, float* x0, cudaStream_t stream);
int launch_finalize(ppgs_engine* e, const float* logits, int ldl, const ForwardPlan& plan,
                    const SeqInfo* seqs_dev, int softmax, float* out, cudaStream_t stream);

// SGEMM of the CUDA-core path: out[M,N] = A[M,K] (row stride lda; rows may overlap, which
// turns strided / 'same' convolutions over time-major activations into GEMMs) * B[N,K]^T
enum { EPI_BIAS = 0, EPI_BIAS_RELU = 1, EPI_BIAS_RES = 2, EPI_IN = 3, EPI_BIAS_GELU = 4 };

struct SgemmArgs {
    const float* A;
    int64_t lda;
    const float* B;   // [N][K]
    const float* bias;
    float* out;
    int64_t ldo;
    int M, N, K;
    const float* res;      // EPI_BIAS_RES: [M][N]
    const float* pe;       // EPI_IN: [max_len][N]
    const SeqInfo* seqs;   // EPI_IN
    const int* tile_seq;
};

int launch_sgemm_any(ppgs_engine* e, const char* name, int epi, const SgemmArgs& a,
                     cudaStream_t stream);
int launch_attention_fp32_any(ppgs_engine* e, int head_dim, const float* qkv, int H, int heads,
                              int max_pitch, int nseq, const SeqInfo* seqs, int causal,
                              float* out, cudaStream_t stream);

}  // namespace ppgs
