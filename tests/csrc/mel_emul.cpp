// CPU replay of the mel kernel's per-frame arithmetic (ppgs_b200/csrc/mel_math.cuh)
// for `-m "not gpu"` tests: a warp is a loop over 32 lanes, shared memory is an
// array.  TEST HARNESS ONLY — never linked into libppgs_b200.so.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (tests/conftest.py).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../ppgs_b200/csrc/mel_math.cuh"

using namespace ppgs;

static uint16_t float_to_half_rn(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));
    if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        int shift = 14 - exp;
        uint32_t half_man = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half_man & 1))) half_man++;
        return (uint16_t)(sign | half_man);
    }
    uint32_t half = sign | ((uint32_t)exp << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++;
    return (uint16_t)half;
}

static float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu, x;
    if (exp == 0) {
        if (man == 0) x = sign;
        else {
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            x = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) x = sign | 0x7f800000u | (man << 13);
    else x = sign | ((exp - 15 + 127) << 23) | (man << 13);
    float f;
    memcpy(&f, &x, 4);
    return f;
}

// Pieces the filterbank is cut into for `basis` ([80][513]); -1 when it does not fit the kernel.
extern "C" int mel_filterbank_rounds(const float* basis) {
    std::vector<FilterbankLayout> layout(1);
    return build_filterbank_layout(basis, layout.data()) ? layout[0].rounds : -1;
}

extern "C" int mel_emul(const float* audio, int batch, long samples, const float* window,
                        const float* tw512f, const float* tw1024f, const float* basis,
                        uint16_t* mel_out) {
    std::vector<FilterbankLayout> layout(1);
    if (!build_filterbank_layout(basis, layout.data())) return -1;
    const FilterbankLayout& fb = layout[0];
    std::vector<float> partial(kFbMaxRounds * 32);
    const long frames = samples / kHop;
    const cf* tw512 = reinterpret_cast<const cf*>(tw512f);
    const cf* tw1024 = reinterpret_cast<const cf*>(tw1024f);
    std::vector<float> frame(kNfft);
    std::vector<cf> zb(kZPad), tmp(kHalf);
    std::vector<float> spec(kFbPartialOffset, 0.f);   // the kernel's slots 513..519 hold finite leftovers: weight 0
    for (int b = 0; b < batch; ++b) {
        const float* row = audio + (long)b * samples;
        for (long f = 0; f < frames; ++f) {
            for (int i = 0; i < kNfft; ++i) {
                long src = f * kHop - kReflect + i;
                if (src < 0) src = -src;
                if (src >= samples) src = 2 * (samples - 1) - src;
                frame[i] = row[src];
            }
            cf v[8];
            // pass 0
            for (int j = 0; j < 64; ++j) {
                for (int r = 0; r < 8; ++r) {
                    int n = j + 64 * r;
                    v[r] = {frame[2 * n] * window[2 * n], frame[2 * n + 1] * window[2 * n + 1]};
                }
                dft8(v);
                int base = stockham_store_base<0>(j);
                for (int r = 0; r < 8; ++r) tmp[base + r] = v[r];
            }
            for (int i = 0; i < kHalf; ++i) zb[zpad(i)] = tmp[i];
            // pass 1
            for (int j = 0; j < 64; ++j) {
                for (int r = 0; r < 8; ++r) v[r] = zb[zpad(j + 64 * r)];
                stockham_twiddle<1>(v, j, tw512);
                dft8(v);
                int base = stockham_store_base<1>(j);
                for (int r = 0; r < 8; ++r) tmp[base + 8 * r] = v[r];
            }
            for (int i = 0; i < kHalf; ++i) zb[zpad(i)] = tmp[i];
            // pass 2
            for (int j = 0; j < 64; ++j) {
                for (int r = 0; r < 8; ++r) v[r] = zb[zpad(j + 64 * r)];
                stockham_twiddle<2>(v, j, tw512);
                dft8(v);
                for (int r = 0; r < 8; ++r) tmp[j + 64 * r] = v[r];
            }
            for (int i = 0; i < kHalf; ++i) zb[zpad(i)] = tmp[i];
            for (int k = 0; k < kBins; ++k)
                spec[k] = half_to_float(float_to_half_rn(sqrtf(bin_power(zb.data(), k, tw1024))));
            for (int r = 0; r < fb.rounds; ++r)
                for (int lane = 0; lane < 32; ++lane)
                    partial[r * 32 + lane] = filterbank_piece(fb.w + r * kFbPiece * 32, fb.base + r * 32,
                                                              lane, spec.data());
            for (int m = 0; m < kMels; ++m) {
                const float acc = filterbank_band(partial.data(), fb.band_slot[m], fb.band_pieces[m]);
                mel_out[((long)b * kMels + m) * frames + f] =
                    float_to_half_rn(logf(fmaxf(acc, 1e-5f)));
            }
        }
    }
    return 0;
}
