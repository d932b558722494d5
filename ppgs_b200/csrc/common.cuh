// Shared declarations of the ppgs_b200 CUDA library (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/ppgs_b200.h"

namespace ppgs {

constexpr int kHopSamples = 160;   // ppgs.HOPSIZE
constexpr int kMelChannels = 80;   // ppgs.NUM_MELS

void set_error(const char* fmt, ...);

// Function attributes (opt-in dynamic shared memory) are per DEVICE: one flag per device
// and kernel instantiation, so a process driving several GPUs configures each of them.
struct PerDeviceOnce {
    bool done[64] = {};
    bool first(int device) {
        if (device < 0 || device >= 64) return true;
        const bool was = done[device];
        done[device] = true;
        return !was;
    }
};

#define PPGS_CUDA(expr)                                                          \
    do {                                                                         \
        cudaError_t err__ = (expr);                                              \
        if (err__ != cudaSuccess) {                                              \
            ::ppgs::set_error("%s failed: %s (%s:%d)", #expr,                    \
                              cudaGetErrorString(err__), __FILE__, __LINE__);    \
            return PPGS_E_CUDA;                                                  \
        }                                                                        \
    } while (0)

#define PPGS_CHECK(expr)                    \
    do {                                    \
        int rc__ = (expr);                  \
        if (rc__ != PPGS_OK) return rc__;   \
    } while (0)

// One folded sequence = (chunk i, utterance b) of ppgs/model/transformer.py:56-63,
// or the whole utterance when un-chunked.  Rows of every activation matrix are
// time-major: sequence s owns rows [row0, row0 + pitch), of which the first
// `tensor_len` are real frames of the chunk tensor and the rest are zero padding.
struct SeqInfo {
    int32_t row0;         // first activation row
    int32_t tensor_len;   // Tc: frames of the chunk tensor (<= pitch - 2)
    int32_t valid_len;    // chunk_lengths[b]: keys / frames that are not masked
    int32_t batch;        // source utterance
    int32_t src_start;    // frame of the utterance at local t=0 (may be negative:
                          // replicate padding clamps to frame 0)
    int32_t keep_begin;   // local frames [keep_begin, keep_end) are emitted ...
    int32_t keep_end;
    int32_t out_start;    // ... to output frames out_start + (t - keep_begin)
};

struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
};

struct MelTables {
    float* window = nullptr;      // [1024] periodic hann
    float2* tw512 = nullptr;      // [512] exp(-2 pi i m / 512)
    float2* tw1024 = nullptr;     // [513] exp(-2 pi i k / 1024)
    // filterbank cut into pieces of kFbPiece bins (mel_math.cuh: FilterbankLayout)
    float* fb_w = nullptr;          // [rounds][kFbPiece][32]
    int32_t* fb_base = nullptr;     // [rounds][32]
    int32_t* band_slot = nullptr;   // [80]
    int32_t* band_pieces = nullptr; // [80]
    int fb_rounds = 0;
};

// Split-fp16 operand planes: plane 0 = hi, plane 1 = lo (rows stacked).
struct PlanePair {
    __half* data = nullptr;   // [2][rows][cols]
    int rows = 0, cols = 0;
    float inv_scale = 1.f;    // multiply the accumulator by this (power of two)
};

// One weight matrix of the tensor-core path: split-fp16 planes
// [2][taps][N][C] (plane 0 = fp16(w*s), plane 1 = fp16(w*s - hi)), the device
// scalar 1/s (s = power of two that moves max|w| near 2^10, so the lo plane stays
// in fp16's normal range), and the TMA descriptor over the planes.
struct TcWeight {
    __half* planes = nullptr;
    const float* inv_scale = nullptr;
    int N = 0, C = 0, taps = 1;
    // TMA descriptors by W-tile rows (256 / 128 = CTA-pair halves / 64 / 32 = fused-FFN
    // halves); index [0] loads the hi plane only (PPGS_PRECISION_F16), [1] both planes
    struct Maps {
        CUtensorMap bn256, bn128, bn64, bn32;
    } maps[2];
};

struct TcLayer {
    TcWeight in_w, out_w, l1_w, l2_w;
};

struct LayerWeights {
    float *in_w, *in_b, *out_w, *out_b, *l1_w, *l1_b, *l2_w, *l2_b;
    float *n1_w, *n1_b, *n2_w, *n2_b;
};

}  // namespace ppgs

namespace ppgs {
struct W2v2Weights;
}

struct ppgs_engine {
    ppgs_model_config cfg;
    int device = 0;
    int precision = PPGS_PRECISION_FP32;
    bool precision_chosen = false;   // false: finalize picks the tensor-core parity mode when the shape allows
    bool finalized = false;
    int sm_count = 148;

    // fp32 weights by reference state-dict key, held on the host until finalize
    std::map<std::string, ppgs::HostTensor> weights;

    // packed blob (one allocation; broadcastable)
    void* blob = nullptr;
    size_t blob_bytes = 0;
    float* conv_in_w = nullptr;    // [H][5*C]   k = tap*C + c
    float* conv_out_w = nullptr;   // [O][5*H]
    float* pe = nullptr;           // [max_len][H]
    std::vector<ppgs::LayerWeights> layers;
    float *conv_in_b = nullptr, *conv_out_b = nullptr;

    // tensor-core path (same blob)
    float* tc_scales = nullptr;
    ppgs::TcWeight tc_conv_in, tc_conv_out;
    std::vector<ppgs::TcLayer> tc_layers;
    bool tc_maps_ready = false;
    int* status_dev = nullptr;   // kernels report barrier time-outs here
    int attention_impl = 1;      // 1 = tcgen05 kernel, 0 = CUDA-core kernel (validation)
    // attention operand planes of the PPG Transformer (not the wav2vec2 encoder): Q / K and the
    // softmax numerators P enter their MMAs as one fp16 plane (S in one pass, P.V in two; measured
    // +1e-5 on the posteriorgram, profiles/r02_attn_planes.jsonl); 2 = hi + lo like every other operand
    int attn_qk_planes = 1;      // PPGS_B200_ATTN_QK_PLANES
    int l2_hints = 0;            // PPGS_B200_L2_HINTS: evict-first loads of operands that die with the kernel (measured: slightly slower, off)
    int serpentine = 1;          // PPGS_B200_SERPENTINE: consecutive kernels of a forward walk the row tiles in opposite directions
    int attn_reverse = 0;        // direction of the next attention launch (set by the forward)
    int mel_rows = 1;            // PPGS_B200_MEL_ROWS: from_audio's mel kernel writes the input convolution's operand rows itself (no fold pass)
    int qk_gemm_passes = 3;      // PPGS_B200_QK_GEMM_PASSES: MMA passes of the Q / K columns of the QKV GEMM when they are kept as one plane
    int attn_p_planes = 1;       // PPGS_B200_ATTN_P_PLANES
    int attn_dual = 1;           // head_dim 128: two query tiles per CTA (attention_dual_tc.cu; PPGS_B200_ATTN_DUAL)
    int gemm_pair = 1;           // 1 = CTA-pair (cta_group::2) GEMMs at BN = 256
    int proj_ln = 1;             // 1 = hidden-256 projections + residual + LayerNorm through proj_ln_kernel (ffn_tc.cu);
                                 // 0 (PPGS_B200_PROJ_LN=0) = gemm_tc's ResLN epilogue
    int fused_ffn = 1;           // 1 = one fused kernel for linear1 + ReLU + linear2 + residual + LN (ffn_tc.cu): the
                                 // hidden activation never leaves the SM (used from sm_count / 8 row-tile pairs up;
                                 // 2 = always); 0 (PPGS_B200_FUSED_FFN=0) = two GEMMs
    unsigned long long* trace_dev = nullptr;   // [8 kernel kinds][8] cycle counters (PPGS_B200_TRACE=1)

    // wav2vec2-base front-end of the `w2v2fb` representation (optional)
    std::map<std::string, ppgs::HostTensor> w2v2_host;
    ppgs::W2v2Weights* w2v2 = nullptr;

    ppgs::MelTables mel;
    std::vector<float> host_window;   // optional override ("frontend.window")

    // grow-only workspace
    void* workspace = nullptr;
    size_t workspace_bytes = 0;
    // pinned staging for per-call tables
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    // device staging for the device-buffer fused entry point (mel features)
    void* io_dev = nullptr;
    size_t io_dev_bytes = 0;
    // host-buffer entry points: two request slots so that the H2D copy of request
    // i+1 and the D2H copy of request i-1 overlap the kernels of request i
    struct HostSlot {
        void* dev = nullptr;
        size_t bytes = 0;
        cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
        bool busy = false;
    };
    HostSlot host_slots[2];
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    uint64_t submitted = 0;
    // plan tables already resident on the device (skip the upload when a call
    // repeats the previous call's shapes and lengths)
    std::vector<ppgs::SeqInfo> cached_plan;
    void* cached_plan_dev = nullptr;
    cudaEvent_t plan_uploaded = nullptr;

    int64_t launches = 0;

    // CUDA-graph cache of the fused from_audio forward (engine.cu): steady-state loops replay
    // the 29 launches as one graph.  An entry is valid while `graph_generation` is unchanged
    // (weights, precision, workspace block, profiling) and may be replayed only while the
    // workspace still holds the plan tables of its plan (`ws_owner`; every other workspace
    // user resets it through ensure_workspace).
    struct GraphEntry {
        std::vector<int64_t> key;
        cudaGraphExec_t exec = nullptr;
        uint64_t generation = 0;
        int plan_id = 0;
        int64_t launches = 0;
        uint64_t last_used = 0;
    };
    std::vector<GraphEntry> graphs;
    std::map<std::vector<int64_t>, int> plan_ids;
    std::vector<std::vector<int64_t>> uncapturable;
    uint64_t graph_generation = 1, graph_clock = 0;
    int ws_owner = 0, next_plan_id = 1;
    bool capturing = false;
    int graphs_enabled = 1;          // PPGS_B200_GRAPHS=0 disables
    cudaStream_t capture_stream = nullptr;
    int64_t graph_replays = 0;

    // optional per-kernel timing (ppgs_engine_set_profiling): CUDA events on the
    // launch stream around every launch, accumulated per kernel name
    bool profiling = false;
    struct KernelStat {
        double ms = 0.0;
        int64_t launches = 0;
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    };
    std::map<std::string, KernelStat> stats;
    std::vector<cudaEvent_t> event_pool;
};

namespace ppgs {

int ensure_workspace(ppgs_engine* e, size_t bytes);
float pack_planes(const HostTensor* w, std::vector<__half>& planes);

// Brackets one kernel launch: counts it, and when profiling records CUDA events
// on `stream` around it.
struct LaunchScope {
    ppgs_engine* e;
    cudaStream_t stream;
    ppgs_engine::KernelStat* stat = nullptr;
    cudaEvent_t start = nullptr;
    LaunchScope(ppgs_engine* e, const char* name, cudaStream_t stream);
    ~LaunchScope();
};
int ensure_pinned(ppgs_engine* e, size_t bytes);

// engine.cu, shared with io.cu: mel front-end + transformer on device buffers (`mel` = fp16
// feature staging of batch x 80 x samples/160)
extern "C" int ppgs_detail_from_audio_device(ppgs_engine* e, const float* audio, int batch, int64_t samples,
                                             int64_t stride, const int64_t* lengths, int softmax,
                                             int legacy_mode, float* out, __half* mel, cudaStream_t stream);

// mel.cu
int build_mel_tables(ppgs_engine* e, const float* basis_host /* [80][513] or null */);
int launch_mel(ppgs_engine* e, const float* audio, int batch, int64_t samples,
               int64_t stride, __half* mel, cudaStream_t stream);

// transformer_fp32.cu
struct ForwardPlan {
    std::vector<SeqInfo> seqs;
    int rows = 0;          // total activation rows (multiple of 128)
    int max_pitch = 0;
    int batch = 0, frames = 0;
};
int build_plan(const ppgs_engine* e, int batch, int frames, const int64_t* lengths,
               int legacy_mode, ForwardPlan* plan);
int transformer_forward_fp32(ppgs_engine* e, const __half* features, const ForwardPlan& plan,
                             int softmax, float* out, cudaStream_t stream);

// w2v2_fp32.cu
int w2v2_accepts_key(const std::string& key, const std::vector<int64_t>& shape);   // 1 = ignore
int w2v2_finalize(ppgs_engine* e);
void w2v2_free(ppgs_engine* e);
int w2v2fb_forward(ppgs_engine* e, const float* audio, int batch, int64_t samples, int64_t stride,
                   const int64_t* lengths, __half* out, cudaStream_t stream);

// transformer_tc.cu
int build_weight_map(TcWeight& w);        // TMA descriptors of one packed weight
int ensure_status_word(ppgs_engine* e);
bool tensor_core_shape(const ppgs_model_config& c);   // model shapes the tcgen05 path covers
int build_weight_maps(ppgs_engine* e);
int transformer_forward_tc(ppgs_engine* e, const __half* features, const ForwardPlan& plan,
                           int softmax, float* out, cudaStream_t stream);
// flac.cu
int flac_read_pcm16(const char* path, int16_t* dst, int64_t capacity, int64_t expect_frames);
int launch_mel_rows(ppgs_engine* e, const float* audio, int batch, int64_t samples, int64_t stride,
                    const ForwardPlan& plan, int legacy_mode, __half* x0, const SeqInfo* seqs_dev,
                    cudaStream_t stream);   // mel.cu
int transformer_tc_input_rows(ppgs_engine* e, const ForwardPlan& plan, cudaStream_t stream, __half** x0,
                              const SeqInfo** seqs_dev);
int check_status(ppgs_engine* e, cudaStream_t stream);

}  // namespace ppgs
