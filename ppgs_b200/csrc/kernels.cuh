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

}  // namespace ppgs
