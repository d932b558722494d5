// Per-frame arithmetic of the fused STFT + mel kernel, shared verbatim by the
// CUDA kernel (mel.cu) and by the CPU emulation harness under tests/csrc that
// checks it against the oracle without a GPU.
//
// Replaces ppgs/preprocess/spectrogram.py:14-50 (reflect pad 432, hann-windowed
// 1024-point STFT hop 160, sqrt(re^2+im^2+1e-6), fp16 round) and
// ppgs/preprocess/mel.py:56-76 (80x513 Slaney filterbank, log(clamp(.,1e-5)),
// fp16 round).
//
// One warp transforms one frame: the 1024 real samples are packed into 512
// complex points, transformed by a 3-pass radix-8 Stockham FFT (in place in a
// padded shared-memory buffer, two butterflies per lane per pass), unpacked to
// the 513 one-sided bins, and reduced by the sparse triangular filterbank.
// All functions are written per (lane, butterfly) so that the host harness can
// replay a warp with a plain loop.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define PPGS_HD __host__ __device__ __forceinline__
#else
#define PPGS_HD inline
#endif

namespace ppgs {

constexpr int kHop = 160;
constexpr int kNfft = 1024;
constexpr int kHalf = 512;          // complex FFT length
constexpr int kBins = 513;
constexpr int kMels = 80;
constexpr int kReflect = (kNfft - kHop) / 2;   // 432
constexpr int kZPad = kHalf;                   // complex FFT buffer length (swizzled, no padding)

struct cf {
    float x, y;
};

// Physical slot of complex element i in the shared FFT buffer: an XOR swizzle inside aligned
// blocks of 16 elements (8-byte elements, 16 "double banks" per half-warp wavefront).  Every
// access pattern of the transform is then conflict-free per half-warp: consecutive elements
// (a permutation of one block), the stride-8 scatter of pass 0 (index 8 l + r: the low three
// bits are XORed with l / 2, bit 3 with l >> 3) and the two-groups-64-apart scatter of pass 1
// (bit 3 is XORed with bit 6 of the index, which separates the groups).
PPGS_HD int zpad(int i) { return i ^ (((i >> 4) & 7) | (((i >> 6) & 1) << 3)); }

PPGS_HD cf cadd(cf a, cf b) { return {a.x + b.x, a.y + b.y}; }
PPGS_HD cf csub(cf a, cf b) { return {a.x - b.x, a.y - b.y}; }
PPGS_HD cf cmul(cf a, cf b) {
    return {fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x)};
}
PPGS_HD cf mul_neg_i(cf a) { return {a.y, -a.x}; }

// 4-point forward DFT, natural order in and out.
PPGS_HD void dft4(cf a0, cf a1, cf a2, cf a3, cf& x0, cf& x1, cf& x2, cf& x3) {
    cf t0 = cadd(a0, a2), t2 = csub(a0, a2);
    cf t1 = cadd(a1, a3), t3 = mul_neg_i(csub(a1, a3));
    x0 = cadd(t0, t1);
    x2 = csub(t0, t1);
    x1 = cadd(t2, t3);
    x3 = csub(t2, t3);
}

// 8-point forward DFT (decimation in frequency), natural order in and out.
PPGS_HD void dft8(cf* v) {
    const float c = 0.70710678118654752440f;
    cf u0 = cadd(v[0], v[4]), d0 = csub(v[0], v[4]);
    cf u1 = cadd(v[1], v[5]), d1 = csub(v[1], v[5]);
    cf u2 = cadd(v[2], v[6]), d2 = csub(v[2], v[6]);
    cf u3 = cadd(v[3], v[7]), d3 = csub(v[3], v[7]);
    d1 = {c * (d1.x + d1.y), c * (d1.y - d1.x)};      // * exp(-i pi/4)
    d2 = mul_neg_i(d2);                               // * exp(-i pi/2)
    d3 = {c * (d3.y - d3.x), -c * (d3.x + d3.y)};     // * exp(-3i pi/4)
    dft4(u0, u1, u2, u3, v[0], v[2], v[4], v[6]);
    dft4(d0, d1, d2, d3, v[1], v[3], v[5], v[7]);
}

// Butterfly `j` (0..63) of Stockham pass `pass` (0,1,2 <-> Ns = 1,8,64).
// Load: v[r] = src[j + 64 r] (twiddled), store: dst[expand(j) + r Ns].
template <int PASS>
PPGS_HD void stockham_twiddle(cf* v, int j, const cf* tw512) {
    if (PASS == 0) return;
    constexpr int Ns = (PASS == 1) ? 8 : 64;
    constexpr int step = 512 / (Ns * 8);
    const int k = j & (Ns - 1);
#pragma unroll
    for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], tw512[k * r * step]);
}

// Same twiddles from per-pass tables laid out [r - 1][k] (k = j mod Ns), so that the lanes
// of a warp read consecutive entries (pass 2) or eight consecutive entries broadcast (pass 1)
// instead of the strided tw512[k r step]: table[(r - 1) Ns + k] == tw512[k r step].
template <int PASS>
PPGS_HD void stockham_twiddle_table(cf* v, int j, const cf* table) {
    if (PASS == 0) return;
    constexpr int Ns = (PASS == 1) ? 8 : 64;
    const int k = j & (Ns - 1);
#pragma unroll
    for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], table[(r - 1) * Ns + k]);
}

template <int PASS>
PPGS_HD int stockham_store_base(int j) {
    constexpr int Ns = (PASS == 0) ? 1 : (PASS == 1) ? 8 : 64;
    return (j / Ns) * Ns * 8 + (j & (Ns - 1));
}

template <int PASS>
PPGS_HD int stockham_store_stride() {
    return (PASS == 0) ? 1 : (PASS == 1) ? 8 : 64;
}

// One-sided bin k (0..512) of the real transform from the packed complex FFT Z
// (zb is the padded buffer), then magnitude as the reference computes it, rounded
// through fp16 and returned as float.
PPGS_HD float bin_power(const cf* zb, int k, const cf* tw1024) {
    cf zk = zb[zpad(k & (kHalf - 1))];
    cf zn = zb[zpad((kHalf - k) & (kHalf - 1))];
    cf e = {0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y)};
    cf o = {0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x)};
    cf x = cadd(e, cmul(tw1024[k], o));
    float p = x.x * x.x + x.y * x.y;   // compiled without fma contraction
    return p + 1e-6f;
}

// Bins k and 512 - k (1 <= k <= 255) from the one pair (Z[k], Z[512 - k]) they share: with
// t = W^k O, bin k is E + t and bin 512 - k is conj(E - t) (tw1024[512 - k] == -conj(tw1024[k])
// bit for bit for these k), so both powers equal bin_power()'s to the last bit at half the loads.
PPGS_HD void bin_power_pair(const cf* zb, int k, const cf* tw1024, float& p_lo, float& p_hi) {
    cf zk = zb[zpad(k)];
    cf zn = zb[zpad(kHalf - k)];
    cf e = {0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y)};
    cf o = {0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x)};
    cf t = cmul(tw1024[k], o);
    cf x = cadd(e, t), y = csub(e, t);
    float a = x.x * x.x + x.y * x.y;
    float b = y.x * y.x + y.y * y.y;
    p_lo = a + 1e-6f;
    p_hi = b + 1e-6f;
}

// ---- triangular filterbank as balanced pieces ------------------------------------------------
// Band m is a run of `count` consecutive bins; one lane per band leaves most of the warp idle
// (5 bins at the bottom, 60 at the top).  The runs are cut into pieces of kFbPiece bins, piece s
// goes to (round s / 32, lane s % 32), every lane reduces one piece per round with an ascending-k
// fmaf chain, and band m then adds its pieces' partial sums in order.  kFbPiece is odd so that
// the lanes of a round, whose pieces start kFbPiece bins apart inside a wide band, read 32
// different shared-memory banks.
constexpr int kFbPiece = 7;
constexpr int kFbMaxRounds = 9;
constexpr int kFbPartialOffset = 520;     // floats: partial sums sit behind the 513 + 6 spectrum slots

struct FilterbankLayout {
    int rounds;
    float w[kFbMaxRounds * kFbPiece * 32];     // [round][j][lane], zero beyond a piece's bins
    int32_t base[kFbMaxRounds * 32];           // [round][lane] first bin of the piece
    int32_t band_slot[kMels];                  // first piece of band m
    int32_t band_pieces[kMels];
};

// basis: [kMels][kBins] row-major.  False when the basis needs more pieces than the kernel holds.
inline bool build_filterbank_layout(const float* basis, FilterbankLayout* layout) {
    for (int i = 0; i < kFbMaxRounds * kFbPiece * 32; ++i) layout->w[i] = 0.f;
    for (int i = 0; i < kFbMaxRounds * 32; ++i) layout->base[i] = 0;
    int slot = 0;
    for (int m = 0; m < kMels; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < kBins; ++k)
            if (basis[m * kBins + k] != 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        const int count = first < 0 ? 0 : last - first + 1;
        const int pieces = (count + kFbPiece - 1) / kFbPiece;
        layout->band_slot[m] = slot;
        layout->band_pieces[m] = pieces;
        if (slot + pieces > kFbMaxRounds * 32) return false;
        for (int p = 0; p < pieces; ++p, ++slot) {
            const int round = slot / 32, lane = slot % 32, k0 = first + p * kFbPiece;
            layout->base[round * 32 + lane] = k0;
            for (int j = 0; j < kFbPiece && k0 + j <= last; ++j)
                layout->w[(round * kFbPiece + j) * 32 + lane] = basis[m * kBins + k0 + j];
        }
    }
    layout->rounds = (slot + 31) / 32;
    return true;
}

// One lane's piece of one round; spec[] holds the 513 magnitudes (+ 6 finite slots behind them).
PPGS_HD float filterbank_piece(const float* w_round, const int32_t* base_round, int lane, const float* spec) {
    const int k0 = base_round[lane];
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < kFbPiece; ++j) acc = fmaf(w_round[j * 32 + lane], spec[k0 + j], acc);
    return acc;
}

// Band m from the partial sums of its pieces, in piece order.
PPGS_HD float filterbank_band(const float* partial, int slot, int pieces) {
    float acc = 0.f;
    if (pieces > 0) acc = partial[slot];
    for (int p = 1; p < pieces; ++p) acc = acc + partial[slot + p];
    return acc;
}

}  // namespace ppgs
