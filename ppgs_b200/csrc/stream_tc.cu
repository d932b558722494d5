// Stateful streaming decoder for causal models (config/causal_transformer.py: IS_CAUSAL =
// True).  The reference cannot stream (ppgs/model/transformer.py:65-71 builds
// the square mask per call and carries no state); this is the incremental form
// of its un-chunked causal forward (legacy_mode=True) over a growing utterance.
//
// Semantics.  A session holds `streams` independent utterances; every push may bring a
// different number of frames (including none) to each of them, and streams can be finalised
// and reset one by one while the others keep going.  After pushes totalling L feature frames
// a stream's state equals the reference forward over those L frames:
//   * hidden position p depends on features <= p + 2 (input Conv1d k=5 'same') and, through
//     the causal attention, on hidden positions <= p: final once L >= p + 3;
//   * output frame q depends on hidden q-2 .. q+2 (output Conv1d k=5): final once L >= q + 5.
// push() therefore emits frames [emitted, L - 4) (everything up to L when `final`), each
// exactly the value the reference computes for the whole utterance (algorithmic latency: 4
// frames = 40 ms).
//
// State = the activation matrices of the tensor-core path made persistent: every stream owns
// 512 rows (510 frames + the 2 zero rows the k=5 convolutions read as padding) of the feature
// matrix x0, the residual stream x, and ONE QKV MATRIX PER LAYER — the K / V columns of
// rows < L are the attention cache, read in place by the tcgen05 attention kernel, whose
// 512-key TMEM budget sets the session capacity (the reference's own chunked inference never
// shows the model more than 500 frames either).  A push appends n feature rows and recomputes
// only the 128-row tiles that hold a position >= L_old - 4 (GemmParams::win_*, AttnParams::
// q_first_tile): positions whose look-ahead was incomplete are redone with the new frames,
// earlier tiles keep their cached K / V.  Recomputing a tile is idempotent, so cached rows that
// share a tile with new rows are rewritten with identical values.
#include <algorithm>

#include "attention_tc.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace ppgs {
using namespace tc;

constexpr int kStreamPitch = 512;       // rows per stream
constexpr int kStreamCapacity = 510;    // frames per session
constexpr int kStreamLookahead = 4;     // 2 (input conv) + 2 (output conv)

// (B, C, n) fp16 feature frames -> rows [row0 + t0_b, row0 + t0_b + n_b) of the time-major x0;
// counts[b] = {t0_b, n_b}
__global__ void stream_append_kernel(const __half* __restrict__ feats, int C, int n,
                                     const int2* __restrict__ counts, __half* __restrict__ x0) {
    __shared__ __half tile[32][34];
    const int b = blockIdx.z, f_base = blockIdx.x * 32, c_base = blockIdx.y * 32;
    const int2 at = counts[b];
    if (f_base >= at.y) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int i = ty; i < 32; i += 8) {
        const int c = c_base + i, f = f_base + tx;
        tile[i][tx] = (c < C && f < at.y) ? feats[((int64_t)b * C + c) * n + f] : __float2half_rn(0.f);
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c_base + tx, f = f_base + i;
        if (c < C && f < at.y) x0[((int64_t)b * kStreamPitch + at.x + f) * C + c] = tile[tx][i];
    }
}

// reset of single streams: every region of the stream's state reads as zero afterwards (the
// feature rows and the residual stream must be zero beyond the new length; the K / V caches
// are cleared too so that non-finite values of an earlier utterance cannot leak through the
// masked tail of a key block)
struct ClearRegions {
    uint32_t* base[16];       // region start of stream 0, as 32-bit words
    int64_t words[16];        // words per stream in that region
    int count;
};

__global__ void stream_clear_kernel(const int2* __restrict__ flags, ClearRegions regions) {
    const int b = blockIdx.y, r = blockIdx.z;
    if (!flags[b].x || r >= regions.count) return;
    uint4* dst = reinterpret_cast<uint4*>(regions.base[r] + (int64_t)b * regions.words[r]);
    const int64_t n = regions.words[r] / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = make_uint4(0u, 0u, 0u, 0u);
}

}  // namespace ppgs

using namespace ppgs;

struct ppgs_stream {
    ppgs_engine* e = nullptr;
    int streams = 0;
    std::vector<int> length;     // L_b: feature frames pushed so far
    std::vector<int> emitted;    // output frames returned so far
    std::vector<char> finished;
    void* block = nullptr;       // one allocation
    __half *x0 = nullptr, *xh = nullptr, *att = nullptr, *ff = nullptr;
    std::vector<__half*> qkv;    // per layer
    SeqInfo* seqs_dev = nullptr;
    int* tile_seq_dev = nullptr;
    int2* counts_dev = nullptr;  // per stream {t0, n} of the push / reset flags
    int2* counts_host = nullptr; // pinned staging of the same
    cudaEvent_t counts_used = nullptr;
    size_t zero_bytes = 0;       // prefix of `block` that a full reset clears (x0, x, qkv caches)
};

static int stream_reset_state(ppgs_stream* s, cudaStream_t stream) {
    PPGS_CUDA(cudaMemsetAsync(s->block, 0, s->zero_bytes, stream));
    std::fill(s->length.begin(), s->length.end(), 0);
    std::fill(s->emitted.begin(), s->emitted.end(), 0);
    std::fill(s->finished.begin(), s->finished.end(), 0);
    return PPGS_OK;
}

// pinned staging -> device, after the previous use of the staging buffer has been consumed
static int upload_counts(ppgs_stream* s, cudaStream_t stream) {
    PPGS_CUDA(cudaMemcpyAsync(s->counts_dev, s->counts_host, (size_t)s->streams * sizeof(int2),
                              cudaMemcpyHostToDevice, stream));
    PPGS_CUDA(cudaEventRecord(s->counts_used, stream));
    return PPGS_OK;
}

extern "C" {

int ppgs_stream_capacity(void) { return kStreamCapacity; }

int ppgs_stream_create(ppgs_engine* e, int streams, ppgs_stream** out) {
    if (!e || !out || streams <= 0) {
        set_error("stream_create: bad argument");
        return PPGS_E_INVALID;
    }
    *out = nullptr;
    const ppgs_model_config& c = e->cfg;
    if (!e->finalized) {
        set_error("engine has no weights: call ppgs_engine_finalize first");
        return PPGS_E_STATE;
    }
    if (!c.is_causal) {
        set_error("stream_create: streaming needs a causal model (config/causal_transformer.py: IS_CAUSAL)");
        return PPGS_E_INVALID;
    }
    if (c.hidden_channels != 256 || !tensor_core_shape(c) || c.kernel_size != 5) {
        set_error("stream_create: the streaming decoder covers hidden 256 and kernel 5 (got hidden %d, "
                  "kernel %d)", c.hidden_channels, c.kernel_size);
        return PPGS_E_UNSUPPORTED;
    }
    if (e->precision == PPGS_PRECISION_FP32) {
        set_error("stream_create: the streaming decoder runs the tensor-core path (precision f16x2 / f16)");
        return PPGS_E_UNSUPPORTED;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    PPGS_CUDA(cudaSetDevice(e->device));
    ppgs_stream* s = new ppgs_stream();
    s->e = e;
    s->streams = streams;
    s->length.assign(streams, 0);
    s->emitted.assign(streams, 0);
    s->finished.assign(streams, 0);
    const size_t rows = (size_t)streams * kStreamPitch;
    const int C = c.input_channels, H = c.hidden_channels, F = c.ffn_channels;
    Carver w;
    const size_t o_x0 = w.take(rows * C * 2);
    const size_t o_xh = w.take(2 * rows * H * 2);
    std::vector<size_t> o_qkv(c.num_layers);
    for (int l = 0; l < c.num_layers; ++l) o_qkv[l] = w.take(2 * rows * 3 * H * 2);
    s->zero_bytes = w.off;
    const size_t o_att = w.take(2 * rows * H * 2);
    const size_t o_ff = w.take(2 * rows * F * 2);
    const size_t o_seqs = w.take((size_t)streams * sizeof(SeqInfo));
    const size_t o_tiles = w.take(rows / 128 * 4);
    const size_t o_counts = w.take((size_t)streams * sizeof(int2));
    cudaError_t err = cudaMalloc(&s->block, w.off);
    if (err != cudaSuccess) {
        set_error("stream_create: cudaMalloc of %zu bytes failed: %s", w.off, cudaGetErrorString(err));
        delete s;
        if (prev >= 0) cudaSetDevice(prev);
        return PPGS_E_CUDA;
    }
    char* base = static_cast<char*>(s->block);
    s->x0 = reinterpret_cast<__half*>(base + o_x0);
    s->xh = reinterpret_cast<__half*>(base + o_xh);
    for (int l = 0; l < c.num_layers; ++l) s->qkv.push_back(reinterpret_cast<__half*>(base + o_qkv[l]));
    s->att = reinterpret_cast<__half*>(base + o_att);
    s->ff = reinterpret_cast<__half*>(base + o_ff);
    s->seqs_dev = reinterpret_cast<SeqInfo*>(base + o_seqs);
    s->tile_seq_dev = reinterpret_cast<int*>(base + o_tiles);
    s->counts_dev = reinterpret_cast<int2*>(base + o_counts);
    int rc = PPGS_OK;
    if (cudaHostAlloc(reinterpret_cast<void**>(&s->counts_host), (size_t)streams * sizeof(int2),
                      cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->counts_used, cudaEventDisableTiming) != cudaSuccess) {
        set_error("stream_create: pinned staging allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = PPGS_E_CUDA;
    }
    if (cudaMemset(s->block, 0, w.off) != cudaSuccess) {
        set_error("stream_create: cudaMemset failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = PPGS_E_CUDA;
    }
    if (rc == PPGS_OK) rc = build_weight_maps(e);
    if (prev >= 0) cudaSetDevice(prev);
    if (rc != PPGS_OK) {
        cudaFree(s->block);
        if (s->counts_host) cudaFreeHost(s->counts_host);
        if (s->counts_used) cudaEventDestroy(s->counts_used);
        delete s;
        return rc;
    }
    *out = s;
    return PPGS_OK;
}

void ppgs_stream_destroy(ppgs_stream* s) {
    if (!s) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s->e->device);
    cudaDeviceSynchronize();
    cudaFree(s->block);
    cudaFreeHost(s->counts_host);
    cudaEventDestroy(s->counts_used);
    if (prev >= 0) cudaSetDevice(prev);
    delete s;
}

int ppgs_stream_reset(ppgs_stream* s, void* stream) {
    if (!s) {
        set_error("stream is NULL");
        return PPGS_E_INVALID;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    PPGS_CUDA(cudaSetDevice(s->e->device));
    const int rc = stream_reset_state(s, static_cast<cudaStream_t>(stream));
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

int ppgs_stream_length(const ppgs_stream* s) { return s && !s->length.empty() ? s->length[0] : -1; }
int ppgs_stream_emitted(const ppgs_stream* s) { return s && !s->emitted.empty() ? s->emitted[0] : -1; }

int ppgs_stream_state(const ppgs_stream* s, int32_t* lengths, int32_t* emitted) {
    if (!s) {
        set_error("stream is NULL");
        return PPGS_E_INVALID;
    }
    for (int b = 0; b < s->streams; ++b) {
        if (lengths) lengths[b] = s->length[b];
        if (emitted) emitted[b] = s->emitted[b];
    }
    return PPGS_OK;
}

int ppgs_stream_reset_streams(ppgs_stream* s, const int32_t* flags, void* stream_) {
    if (!s || !flags) {
        set_error("stream_reset_streams: bad argument");
        return PPGS_E_INVALID;
    }
    ppgs_engine* e = s->e;
    int prev = -1;
    cudaGetDevice(&prev);
    PPGS_CUDA(cudaSetDevice(e->device));
    struct Restore {
        int prev;
        ~Restore() {
            if (prev >= 0) cudaSetDevice(prev);
        }
    } restore{prev};
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PPGS_CUDA(cudaEventSynchronize(s->counts_used));
    bool any = false;
    for (int b = 0; b < s->streams; ++b) {
        s->counts_host[b] = make_int2(flags[b] ? 1 : 0, 0);
        if (flags[b]) {
            any = true;
            s->length[b] = s->emitted[b] = 0;
            s->finished[b] = 0;
        }
    }
    if (!any) return PPGS_OK;
    PPGS_CHECK(upload_counts(s, stream));
    const ppgs_model_config& c = e->cfg;
    const int64_t rows = (int64_t)s->streams * kStreamPitch;
    ClearRegions regions;
    regions.count = 0;
    auto add = [&](__half* base, int64_t halves_per_stream) {
        regions.base[regions.count] = reinterpret_cast<uint32_t*>(base);
        regions.words[regions.count] = halves_per_stream / 2;
        regions.count += 1;
    };
    add(s->x0, (int64_t)kStreamPitch * c.input_channels);
    for (int plane = 0; plane < 2; ++plane)
        add(s->xh + plane * rows * c.hidden_channels, (int64_t)kStreamPitch * c.hidden_channels);
    for (size_t layer = 0; layer < s->qkv.size() && regions.count + 2 <= 16; ++layer)
        for (int plane = 0; plane < 2; ++plane)
            add(s->qkv[layer] + plane * rows * 3 * c.hidden_channels,
                (int64_t)kStreamPitch * 3 * c.hidden_channels);
    {
        LaunchScope scope(e, "stream_clear", stream);
        stream_clear_kernel<<<dim3(48, s->streams, regions.count), 256, 0, stream>>>(s->counts_dev, regions);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

int ppgs_stream_push_ragged(ppgs_stream* s, const void* features_dev, int max_frames, const int32_t* frames,
                            const int32_t* final, int softmax, float* out_dev, int out_capacity,
                            int32_t* frames_out, void* stream_) {
    if (!s) {
        set_error("stream is NULL");
        return PPGS_E_INVALID;
    }
    ppgs_engine* e = s->e;
    const ppgs_model_config& c = e->cfg;
    const int B = s->streams;
    if (frames_out)
        for (int b = 0; b < B; ++b) frames_out[b] = 0;
    if (max_frames < 0 || !frames || (max_frames > 0 && !features_dev) || out_capacity < 0 ||
        (out_capacity > 0 && !out_dev)) {
        set_error("stream_push: bad argument");
        return PPGS_E_INVALID;
    }
    if (e->precision == PPGS_PRECISION_FP32) {
        set_error("stream_push: the streaming decoder runs the tensor-core path (precision f16x2 / f16)");
        return PPGS_E_UNSUPPORTED;
    }
    // validate everything before any state changes
    for (int b = 0; b < B; ++b) {
        const int n = frames[b];
        if (n < 0 || n > max_frames) {
            set_error("stream_push: stream %d brings %d frames, the feature tensor holds %d", b, n, max_frames);
            return PPGS_E_INVALID;
        }
        if (s->finished[b] && (n > 0 || (final && final[b]))) {
            set_error("stream_push: stream %d was finalised; reset it first", b);
            return PPGS_E_STATE;
        }
        if (s->length[b] + n > kStreamCapacity) {
            // the positional-encoding limit of the reference is 5000 (transformer.py:103-104); here
            // the cache capacity is the binding one
            set_error("size is too large: a streaming session holds %d frames (stream %d: %d pushed + %d new)",
                      kStreamCapacity, b, s->length[b], n);
            return PPGS_E_TOO_LARGE;
        }
        const int L = s->length[b] + n;
        const bool fin = final && final[b];
        const int keep_end = fin ? L : (L - kStreamLookahead > s->emitted[b] ? L - kStreamLookahead : s->emitted[b]);
        if (keep_end - s->emitted[b] > out_capacity) {
            set_error("stream_push: %d output frames of stream %d do not fit out_capacity %d",
                      keep_end - s->emitted[b], b, out_capacity);
            return PPGS_E_INVALID;
        }
    }
    int prev = -1;
    cudaGetDevice(&prev);
    PPGS_CUDA(cudaSetDevice(e->device));
    struct Restore {
        int prev;
        ~Restore() {
            if (prev >= 0) cudaSetDevice(prev);
        }
    } restore{prev};
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int C = c.input_channels, H = c.hidden_channels, F = c.ffn_channels, O = c.output_channels;
    const int k = c.kernel_size;
    const int rows = B * kStreamPitch;
    const int planes = e->precision == PPGS_PRECISION_F16X2 ? 2 : 1;
    const int tiles_per_seq = kStreamPitch / 128;

    // per-stream bookkeeping: new length, frames to emit, tiles to (re)compute — every position
    // >= max(t0 - 4, 0): incomplete look-ahead, and the rows the output convolution of the
    // first emitted frame reads
    ForwardPlan plan;
    plan.rows = rows;
    plan.max_pitch = kStreamPitch;
    plan.batch = B;
    plan.seqs.resize(B);
    std::vector<int> tile_first(B, 0), tile_end(B, 0);
    PPGS_CUDA(cudaEventSynchronize(s->counts_used));
    int win_tiles = 0, max_new = 0, max_out = 0;
    bool any_work = false;
    for (int b = 0; b < B; ++b) {
        const int n = frames[b], t0 = s->length[b], L = t0 + n;
        const bool fin = final && final[b];
        const int keep_end = fin ? L : (L - kStreamLookahead > s->emitted[b] ? L - kStreamLookahead : s->emitted[b]);
        const int n_out = keep_end - s->emitted[b];
        s->counts_host[b] = make_int2(t0, n);
        SeqInfo& q = plan.seqs[b];
        q.row0 = b * kStreamPitch;
        q.tensor_len = L;
        q.valid_len = L;
        q.batch = b;
        q.keep_begin = s->emitted[b];
        q.keep_end = keep_end;
        q.out_start = 0;
        const bool work = L > 0 && (n > 0 || n_out > 0);
        if (work) {
            const int lookback = t0 - kStreamLookahead > 0 ? t0 - kStreamLookahead : 0;
            const int first_pos = s->emitted[b] < lookback ? s->emitted[b] : lookback;
            tile_first[b] = first_pos / 128;
            tile_end[b] = (L - 1) / 128 + 1;
            win_tiles = std::max(win_tiles, tile_end[b] - tile_first[b]);
            any_work = true;
        }
        max_new = std::max(max_new, n);
        max_out = std::max(max_out, n_out);
        if (frames_out) frames_out[b] = n_out;
        s->length[b] = L;
        s->emitted[b] = keep_end;
        if (fin) s->finished[b] = 1;
    }
    if (max_new > 0) {
        PPGS_CHECK(upload_counts(s, stream));
        dim3 grid((max_new + 31) / 32, (C + 31) / 32, B);
        LaunchScope scope(e, "stream_append", stream);
        stream_append_kernel<<<grid, dim3(32, 8), 0, stream>>>(static_cast<const __half*>(features_dev), C,
                                                                max_frames, s->counts_dev, s->x0);
    }
    PPGS_CUDA(cudaGetLastError());
    if (!any_work) return PPGS_OK;

    // one window size for every stream (the CTA-pair GEMMs take an even number of tiles); a
    // stream that needs fewer tiles recomputes neighbours, which is idempotent
    win_tiles = (win_tiles + 1) & ~1;
    if (win_tiles > tiles_per_seq) win_tiles = tiles_per_seq;
    for (int b = 0; b < B; ++b) {
        int first = tile_first[b];
        if (first + win_tiles > tiles_per_seq) first = tiles_per_seq - win_tiles;
        plan.seqs[b].src_start = first;   // GemmParams::win_per_seq / attention q_first_tile = -1
    }
    PPGS_CHECK(upload_plan(e, plan, s->seqs_dev, s->tile_seq_dev, stream));

    CUtensorMap map_x0, map_x, map_att, map_ff, out_x, out_ff;
    PPGS_CHECK(make_plane_map(&map_x0, s->x0, false, C, rows, 1, 1, C, 0, (uint64_t)rows * C, 128, 1));
    PPGS_CHECK(make_plane_map(&map_x, s->xh, false, H, rows, 1, 2, H, 0, (uint64_t)rows * H, 128, planes));
    PPGS_CHECK(make_plane_map(&map_att, s->att, false, H, rows, 1, 2, H, 0, (uint64_t)rows * H, 128, planes));
    PPGS_CHECK(make_plane_map(&map_ff, s->ff, false, F, rows, 1, 2, F, 0, (uint64_t)rows * F, 128, planes));
    PPGS_CHECK(make_store_map(&out_x, s->xh, H, rows, (uint64_t)rows * H));
    PPGS_CHECK(make_store_map(&out_ff, s->ff, F, rows, (uint64_t)rows * F));
    CUtensorMap map_res;   // residual rows of the LayerNorm epilogues: both planes of x
    PPGS_CHECK(make_plane_map(&map_res, s->xh, false, H, rows, 1, 2, H, 0, (uint64_t)rows * H, 128, 2));

    GemmParams base;
    base.m_tiles = B * win_tiles;
    base.seqs = s->seqs_dev;
    base.tile_seq = s->tile_seq_dev;
    base.status = e->status_dev;
    base.eps = c.layer_norm_eps;
    base.b_planes = planes;
    base.pair = 1;
    base.win_size = win_tiles;
    base.win_stride = tiles_per_seq;
    base.win_per_seq = 1;
    auto wmap = [&](TcWeight& w) -> const CUtensorMap& { return w.maps[planes - 1].bn128; };
    // serpentine order over the (session, window tile) items, as in the batch path
    int direction = 0;
    auto next_direction = [&]() {
        const int d = e->serpentine ? direction : 0;
        direction ^= 1;
        return d;
    };

    {
        GemmParams p = base;
        p.reverse = next_direction();
        p.n_tiles = 1; p.taps = k; p.half = k / 2; p.cblocks = (C + 63) / 64; p.a_planes = 1;
        p.N = H; p.scale = e->tc_conv_in.inv_scale; p.bias = e->conv_in_b; p.pe = e->pe;
        PPGS_CHECK(launch_gemm_tc(e, "tc_conv_in", 256, kEpiConvIn, map_x0, wmap(e->tc_conv_in), &out_x, p, stream));
    }
    for (int layer = 0; layer < c.num_layers; ++layer) {
        const LayerWeights& Lw = e->layers[layer];
        TcLayer& T = e->tc_layers[layer];
        CUtensorMap out_qkv;
        PPGS_CHECK(make_store_map(&out_qkv, s->qkv[layer], 3 * H, rows, (uint64_t)rows * 3 * H));
        {
            GemmParams p = base;
            p.n_tiles = 3 * H / 256; p.cblocks = H / 64; p.a_planes = planes;
            p.N = 3 * H; p.scale = T.in_w.inv_scale; p.bias = Lw.in_b;
            p.reverse = next_direction();
            PPGS_CHECK(launch_gemm_tc(e, "tc_qkv", 256, kEpiPlanes, map_x, wmap(T.in_w), &out_qkv, p, stream));
        }
        e->attn_reverse = next_direction();
        PPGS_CHECK(launch_attention_any(e, s->qkv[layer], s->att, rows, H, c.num_heads, kStreamPitch, B,
                                        s->seqs_dev, 1, planes, stream, -1, win_tiles, e->attn_qk_planes,
                                        e->attn_p_planes));
        // the batch path's fused kernels over the same row-tile window (ffn_tc.cu)
        FfnParams f;
        f.m_tiles = B * win_tiles; f.planes = planes;
        f.win_size = win_tiles; f.win_stride = tiles_per_seq;
        f.eps = c.layer_norm_eps; f.seqs = s->seqs_dev; f.tile_seq = s->tile_seq_dev;
        f.status = e->status_dev; f.trace = nullptr;
        e->attn_reverse = 0;
        f.reverse = next_direction();
        if (e->proj_ln && H == 256) {
            f.num_chunks = H / 64;
            f.scale1 = T.out_w.inv_scale; f.scale2 = T.out_w.inv_scale;
            f.bias1 = Lw.out_b; f.bias2 = Lw.out_b; f.gamma = Lw.n1_w; f.beta = Lw.n1_b;
            PPGS_CHECK(launch_proj_ln(e, "tc_out_proj_ln", map_att, wmap(T.out_w), out_x, map_res, f, stream));
        } else {
            GemmParams p = base;
            p.n_tiles = 1; p.cblocks = H / 64; p.a_planes = planes;
            p.N = H; p.scale = T.out_w.inv_scale; p.bias = Lw.out_b;
            p.residual = s->xh; p.res_ld = H; p.res_plane_stride = (int64_t)rows * H;
            p.gamma = Lw.n1_w; p.beta = Lw.n1_b; p.reverse = f.reverse;
            PPGS_CHECK(launch_gemm_tc(e, "tc_out_proj_ln", 256, kEpiResLN, map_att, wmap(T.out_w), &out_x, p, stream));
        }
        if (e->fused_ffn && H == 256 && F % 128 == 0) {
            f.num_chunks = F / 128;
            f.scale1 = T.l1_w.inv_scale; f.scale2 = T.l2_w.inv_scale;
            f.bias1 = Lw.l1_b; f.bias2 = Lw.l2_b; f.gamma = Lw.n2_w; f.beta = Lw.n2_b;
            f.reverse = next_direction();
            PPGS_CHECK(launch_ffn_fused(e, map_x, T.l1_w.maps[planes - 1].bn64, T.l2_w.maps[planes - 1].bn128, out_x,
                                        map_res, f, stream));
            continue;
        }
        {
            GemmParams p = base;
            p.n_tiles = F / 256; p.cblocks = H / 64; p.a_planes = planes;
            p.N = F; p.scale = T.l1_w.inv_scale; p.bias = Lw.l1_b; p.relu = 1;
            p.reverse = next_direction();
            PPGS_CHECK(launch_gemm_tc(e, "tc_ffn1", 256, kEpiPlanes, map_x, wmap(T.l1_w), &out_ff, p, stream));
        }
        {
            GemmParams p = base;
            p.n_tiles = 1; p.cblocks = F / 64; p.a_planes = planes;
            p.N = H; p.scale = T.l2_w.inv_scale; p.bias = Lw.l2_b;
            p.residual = s->xh; p.res_ld = H; p.res_plane_stride = (int64_t)rows * H;
            p.gamma = Lw.n2_w; p.beta = Lw.n2_b; p.reverse = next_direction();
            PPGS_CHECK(launch_gemm_tc(e, "tc_ffn2_ln", 256, kEpiResLN, map_ff, wmap(T.l2_w), &out_x, p, stream));
        }
    }
    if (max_out > 0) {
        GemmParams p = base;
        p.pair = 0;                       // the BN = 64 kernel is single-CTA
        p.n_tiles = 1; p.taps = k; p.half = k / 2; p.cblocks = H / 64; p.a_planes = planes;
        p.N = O; p.O = O; p.scale = e->tc_conv_out.inv_scale; p.bias = e->conv_out_b;
        p.ppg = out_dev; p.T = out_capacity; p.softmax = softmax; p.reverse = next_direction();
        PPGS_CHECK(launch_gemm_tc(e, "tc_conv_out_softmax", 64, kEpiConvOut, map_x,
                                  e->tc_conv_out.maps[planes - 1].bn64, nullptr, p, stream));
    }
    return PPGS_OK;
}

int ppgs_stream_push(ppgs_stream* s, const void* features_dev, int frames, int final, int softmax,
                     float* out_dev, int out_capacity, int* frames_out, void* stream) {
    if (!s) {
        set_error("stream is NULL");
        return PPGS_E_INVALID;
    }
    if (frames_out) *frames_out = 0;
    if (frames < 0) {
        set_error("stream_push: bad argument");
        return PPGS_E_INVALID;
    }
    std::vector<int32_t> counts(s->streams, frames), finals(s->streams, final ? 1 : 0), produced(s->streams, 0);
    PPGS_CHECK(ppgs_stream_push_ragged(s, features_dev, frames, counts.data(), finals.data(), softmax, out_dev,
                                       out_capacity, produced.data(), stream));
    if (frames_out) *frames_out = produced[0];
    return PPGS_OK;
}

}  // extern "C"
