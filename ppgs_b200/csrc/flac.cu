// FLAC ingest for ppgs.load.audio (ppgs/load.py:17-30 hands whatever the corpus holds to
// torchaudio.load; LibriSpeech-style corpora are FLAC, and torchaudio cannot decode anything in
// this image).  Host code only: a bit-exact decoder of the FLAC format (all subframe
// types, both Rice methods, escaped partitions, wasted bits, the three stereo decorrelations,
// 4..32 bits per sample, up to 8 channels, fixed and variable block size) that VERIFIES what it
// decodes: the CRC-8 of every frame header, the CRC-16 of every frame and, when STREAMINFO
// carries one, the MD5 of the whole decoded signal.  Output is torchaudio's normalisation,
// sample / 2^(bits-1) as fp32, channel-major.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdint>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace ppgs {
namespace {

// ---- MD5 (RFC 1321) of the decoded signal, as the encoder computed it ----------------------
struct Md5 {
    uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
    uint64_t bytes = 0;
    uint8_t block[64];
    size_t fill = 0;

    static uint32_t rotl(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }

    void compress(const uint8_t* p) {
        uint32_t m[16];
        memcpy(m, p, 64);                        // little-endian host (x86-64 / aarch64)
        uint32_t A = a, B = b, C = c, D = d;
#define PPGS_MD5_STEP(f, w, x, y, z, g, k, s) \
    w += f(x, y, z) + m[g] + k;               \
    w = rotl(w, s) + x;
#define PPGS_MD5_F(x, y, z) (z ^ (x & (y ^ z)))
#define PPGS_MD5_G(x, y, z) (y ^ (z & (x ^ y)))
#define PPGS_MD5_H(x, y, z) (x ^ y ^ z)
#define PPGS_MD5_I(x, y, z) (y ^ (x | ~z))
        PPGS_MD5_STEP(PPGS_MD5_F, A, B, C, D, 0, 0xd76aa478u, 7)
        PPGS_MD5_STEP(PPGS_MD5_F, D, A, B, C, 1, 0xe8c7b756u, 12)
        PPGS_MD5_STEP(PPGS_MD5_F, C, D, A, B, 2, 0x242070dbu, 17)
        PPGS_MD5_STEP(PPGS_MD5_F, B, C, D, A, 3, 0xc1bdceeeu, 22)
        PPGS_MD5_STEP(PPGS_MD5_F, A, B, C, D, 4, 0xf57c0fafu, 7)
        PPGS_MD5_STEP(PPGS_MD5_F, D, A, B, C, 5, 0x4787c62au, 12)
        PPGS_MD5_STEP(PPGS_MD5_F, C, D, A, B, 6, 0xa8304613u, 17)
        PPGS_MD5_STEP(PPGS_MD5_F, B, C, D, A, 7, 0xfd469501u, 22)
        PPGS_MD5_STEP(PPGS_MD5_F, A, B, C, D, 8, 0x698098d8u, 7)
        PPGS_MD5_STEP(PPGS_MD5_F, D, A, B, C, 9, 0x8b44f7afu, 12)
        PPGS_MD5_STEP(PPGS_MD5_F, C, D, A, B, 10, 0xffff5bb1u, 17)
        PPGS_MD5_STEP(PPGS_MD5_F, B, C, D, A, 11, 0x895cd7beu, 22)
        PPGS_MD5_STEP(PPGS_MD5_F, A, B, C, D, 12, 0x6b901122u, 7)
        PPGS_MD5_STEP(PPGS_MD5_F, D, A, B, C, 13, 0xfd987193u, 12)
        PPGS_MD5_STEP(PPGS_MD5_F, C, D, A, B, 14, 0xa679438eu, 17)
        PPGS_MD5_STEP(PPGS_MD5_F, B, C, D, A, 15, 0x49b40821u, 22)
        PPGS_MD5_STEP(PPGS_MD5_G, A, B, C, D, 1, 0xf61e2562u, 5)
        PPGS_MD5_STEP(PPGS_MD5_G, D, A, B, C, 6, 0xc040b340u, 9)
        PPGS_MD5_STEP(PPGS_MD5_G, C, D, A, B, 11, 0x265e5a51u, 14)
        PPGS_MD5_STEP(PPGS_MD5_G, B, C, D, A, 0, 0xe9b6c7aau, 20)
        PPGS_MD5_STEP(PPGS_MD5_G, A, B, C, D, 5, 0xd62f105du, 5)
        PPGS_MD5_STEP(PPGS_MD5_G, D, A, B, C, 10, 0x02441453u, 9)
        PPGS_MD5_STEP(PPGS_MD5_G, C, D, A, B, 15, 0xd8a1e681u, 14)
        PPGS_MD5_STEP(PPGS_MD5_G, B, C, D, A, 4, 0xe7d3fbc8u, 20)
        PPGS_MD5_STEP(PPGS_MD5_G, A, B, C, D, 9, 0x21e1cde6u, 5)
        PPGS_MD5_STEP(PPGS_MD5_G, D, A, B, C, 14, 0xc33707d6u, 9)
        PPGS_MD5_STEP(PPGS_MD5_G, C, D, A, B, 3, 0xf4d50d87u, 14)
        PPGS_MD5_STEP(PPGS_MD5_G, B, C, D, A, 8, 0x455a14edu, 20)
        PPGS_MD5_STEP(PPGS_MD5_G, A, B, C, D, 13, 0xa9e3e905u, 5)
        PPGS_MD5_STEP(PPGS_MD5_G, D, A, B, C, 2, 0xfcefa3f8u, 9)
        PPGS_MD5_STEP(PPGS_MD5_G, C, D, A, B, 7, 0x676f02d9u, 14)
        PPGS_MD5_STEP(PPGS_MD5_G, B, C, D, A, 12, 0x8d2a4c8au, 20)
        PPGS_MD5_STEP(PPGS_MD5_H, A, B, C, D, 5, 0xfffa3942u, 4)
        PPGS_MD5_STEP(PPGS_MD5_H, D, A, B, C, 8, 0x8771f681u, 11)
        PPGS_MD5_STEP(PPGS_MD5_H, C, D, A, B, 11, 0x6d9d6122u, 16)
        PPGS_MD5_STEP(PPGS_MD5_H, B, C, D, A, 14, 0xfde5380cu, 23)
        PPGS_MD5_STEP(PPGS_MD5_H, A, B, C, D, 1, 0xa4beea44u, 4)
        PPGS_MD5_STEP(PPGS_MD5_H, D, A, B, C, 4, 0x4bdecfa9u, 11)
        PPGS_MD5_STEP(PPGS_MD5_H, C, D, A, B, 7, 0xf6bb4b60u, 16)
        PPGS_MD5_STEP(PPGS_MD5_H, B, C, D, A, 10, 0xbebfbc70u, 23)
        PPGS_MD5_STEP(PPGS_MD5_H, A, B, C, D, 13, 0x289b7ec6u, 4)
        PPGS_MD5_STEP(PPGS_MD5_H, D, A, B, C, 0, 0xeaa127fau, 11)
        PPGS_MD5_STEP(PPGS_MD5_H, C, D, A, B, 3, 0xd4ef3085u, 16)
        PPGS_MD5_STEP(PPGS_MD5_H, B, C, D, A, 6, 0x04881d05u, 23)
        PPGS_MD5_STEP(PPGS_MD5_H, A, B, C, D, 9, 0xd9d4d039u, 4)
        PPGS_MD5_STEP(PPGS_MD5_H, D, A, B, C, 12, 0xe6db99e5u, 11)
        PPGS_MD5_STEP(PPGS_MD5_H, C, D, A, B, 15, 0x1fa27cf8u, 16)
        PPGS_MD5_STEP(PPGS_MD5_H, B, C, D, A, 2, 0xc4ac5665u, 23)
        PPGS_MD5_STEP(PPGS_MD5_I, A, B, C, D, 0, 0xf4292244u, 6)
        PPGS_MD5_STEP(PPGS_MD5_I, D, A, B, C, 7, 0x432aff97u, 10)
        PPGS_MD5_STEP(PPGS_MD5_I, C, D, A, B, 14, 0xab9423a7u, 15)
        PPGS_MD5_STEP(PPGS_MD5_I, B, C, D, A, 5, 0xfc93a039u, 21)
        PPGS_MD5_STEP(PPGS_MD5_I, A, B, C, D, 12, 0x655b59c3u, 6)
        PPGS_MD5_STEP(PPGS_MD5_I, D, A, B, C, 3, 0x8f0ccc92u, 10)
        PPGS_MD5_STEP(PPGS_MD5_I, C, D, A, B, 10, 0xffeff47du, 15)
        PPGS_MD5_STEP(PPGS_MD5_I, B, C, D, A, 1, 0x85845dd1u, 21)
        PPGS_MD5_STEP(PPGS_MD5_I, A, B, C, D, 8, 0x6fa87e4fu, 6)
        PPGS_MD5_STEP(PPGS_MD5_I, D, A, B, C, 15, 0xfe2ce6e0u, 10)
        PPGS_MD5_STEP(PPGS_MD5_I, C, D, A, B, 6, 0xa3014314u, 15)
        PPGS_MD5_STEP(PPGS_MD5_I, B, C, D, A, 13, 0x4e0811a1u, 21)
        PPGS_MD5_STEP(PPGS_MD5_I, A, B, C, D, 4, 0xf7537e82u, 6)
        PPGS_MD5_STEP(PPGS_MD5_I, D, A, B, C, 11, 0xbd3af235u, 10)
        PPGS_MD5_STEP(PPGS_MD5_I, C, D, A, B, 2, 0x2ad7d2bbu, 15)
        PPGS_MD5_STEP(PPGS_MD5_I, B, C, D, A, 9, 0xeb86d391u, 21)
#undef PPGS_MD5_STEP
#undef PPGS_MD5_F
#undef PPGS_MD5_G
#undef PPGS_MD5_H
#undef PPGS_MD5_I
        a += A;
        b += B;
        c += C;
        d += D;
    }

    void update(const uint8_t* p, size_t n) {
        bytes += n;
        if (fill == 0)
            for (; n >= 64; p += 64, n -= 64) compress(p);
        while (n) {
            const size_t take = n < 64 - fill ? n : 64 - fill;
            memcpy(block + fill, p, take);
            fill += take;
            p += take;
            n -= take;
            if (fill == 64) {
                compress(block);
                fill = 0;
            }
        }
    }

    void finish(uint8_t out[16]) {
        const uint64_t bit_count = bytes * 8;
        const uint8_t one = 0x80, zero = 0;
        update(&one, 1);
        while (fill != 56) update(&zero, 1);
        uint8_t len[8];
        for (int i = 0; i < 8; ++i) len[i] = (uint8_t)(bit_count >> (8 * i));
        update(len, 8);
        const uint32_t words[4] = {a, b, c, d};
        for (int i = 0; i < 16; ++i) out[i] = (uint8_t)(words[i / 4] >> (8 * (i % 4)));
    }
};

// ---- CRCs of the frame layer ------------------------------------------------------------------
struct CrcTables {
    uint8_t crc8[256];
    uint16_t crc16[8][256];     // slice-by-8: crc16[k][x] = CRC of byte x followed by k zero bytes
    CrcTables() {
        for (int i = 0; i < 256; ++i) {
            uint8_t c8 = (uint8_t)i;
            uint16_t c16 = (uint16_t)(i << 8);
            for (int bit = 0; bit < 8; ++bit) {
                c8 = (uint8_t)((c8 << 1) ^ ((c8 & 0x80) ? 0x07 : 0));               // x^8 + x^2 + x + 1
                c16 = (uint16_t)((c16 << 1) ^ ((c16 & 0x8000) ? 0x8005 : 0));       // x^16 + x^15 + x^2 + 1
            }
            crc8[i] = c8;
            crc16[0][i] = c16;
        }
        for (int k = 1; k < 8; ++k)
            for (int i = 0; i < 256; ++i)
                crc16[k][i] = (uint16_t)((crc16[k - 1][i] << 8) ^ crc16[0][crc16[k - 1][i] >> 8]);
    }
};
const CrcTables& crc_tables() {
    static const CrcTables tables;
    return tables;
}
uint8_t crc8(const uint8_t* p, size_t n) {
    const CrcTables& t = crc_tables();
    uint8_t c = 0;
    for (size_t i = 0; i < n; ++i) c = t.crc8[c ^ p[i]];
    return c;
}
uint16_t crc16(const uint8_t* p, size_t n) {
    const CrcTables& t = crc_tables();
    uint16_t c = 0;
    for (; n >= 8; p += 8, n -= 8)
        c = (uint16_t)(t.crc16[7][p[0] ^ (c >> 8)] ^ t.crc16[6][p[1] ^ (c & 0xff)] ^ t.crc16[5][p[2]] ^
                       t.crc16[4][p[3]] ^ t.crc16[3][p[4]] ^ t.crc16[2][p[5]] ^ t.crc16[1][p[6]] ^
                       t.crc16[0][p[7]]);
    for (size_t i = 0; i < n; ++i) c = (uint16_t)((c << 8) ^ t.crc16[0][(c >> 8) ^ p[i]]);
    return c;
}

// ---- MSB-first bit reader over the file image ------------------------------------------------
// The image it reads must be followed by >= 8 readable zero bytes (load_file pads).
struct BitReader {
    const uint8_t* data;
    size_t size;
    size_t next;         // next byte to load
    uint64_t window = 0; // valid bits are the top `bits`
    int bits = 0;

    BitReader(const uint8_t* d, size_t n, size_t start) : data(d), size(n), next(start) {}

    void refill() {      // >= 33 valid bits afterwards
        if (bits <= 32) {
            uint32_t word = 0;
            if (next + 4 <= size + 8) memcpy(&word, data + next, 4);
            window |= (uint64_t)__builtin_bswap32(word) << (32 - bits);
            next += 4;
            bits += 32;
        }
    }
    uint32_t read(int count) {                    // 0..32 bits
        if (count == 0) return 0;
        refill();
        const uint32_t value = (uint32_t)(window >> (64 - count));
        window <<= count;
        bits -= count;
        return value;
    }
    int64_t read_signed(int count) {              // 0..33 bits, two's complement
        if (count == 0) return 0;
        uint64_t value;
        if (count > 32) {
            value = (uint64_t)read(count - 32) << 32;
            value |= read(32);
        } else {
            value = read(count);
        }
        return (int64_t)(value << (64 - count)) >> (64 - count);
    }
    bool read_unary(uint32_t* zeros) {            // number of 0 bits before the next 1
        uint32_t count = 0;
        for (;;) {
            refill();
            if (window == 0) {
                count += bits;
                bits = 0;
                if (next >= size + 8) return false;
                continue;
            }
            const int lead = __builtin_clzll(window);
            count += lead;
            window <<= lead;
            window <<= 1;                          // two steps: lead + 1 may be 64
            bits -= lead + 1;
            *zeros = count;
            return true;
        }
    }
    // One Rice-coded residual with parameter k (< 31): unary quotient, k-bit remainder, zig-zag.
    bool read_rice(int k, int64_t* value) {
        refill();
        uint32_t quotient;
        if (window >> 32) {                        // the terminating 1 is within the next 32 bits
            const int lead = __builtin_clzll(window);
            window <<= lead + 1;
            bits -= lead + 1;
            quotient = (uint32_t)lead;
        } else if (!read_unary(&quotient)) {
            return false;
        }
        const uint64_t folded = ((uint64_t)quotient << k) | read(k);
        *value = (int64_t)(folded >> 1) ^ -(int64_t)(folded & 1);
        return true;
    }
    size_t bit_position() const { return next * 8 - bits; }
    bool overrun() const { return bit_position() > size * 8; }
    void align() {
        const int drop = bits & 7;
        window <<= drop;
        bits -= drop;
    }
    size_t byte_position() const { return bit_position() / 8; }
};

struct StreamInfo {
    int min_block = 0, max_block = 0;
    int sample_rate = 0, channels = 0, bits = 0;
    int64_t total = 0;            // 0: not recorded
    uint8_t md5[16] = {0};
    bool has_md5 = false;
    size_t audio_offset = 0;
};

constexpr size_t kImagePad = 16;

// The file followed by kImagePad zero bytes (image->size() - kImagePad is the file size).
bool load_file(const char* path, std::vector<uint8_t>* image) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) {
        set_error("%s: cannot open: %s", path, strerror(errno));
        return false;
    }
    struct stat st;
    if (fstat(fd, &st) != 0) {
        set_error("%s: cannot stat: %s", path, strerror(errno));
        close(fd);
        return false;
    }
    const size_t size = (size_t)st.st_size;
    image->assign(size + kImagePad, 0);          // zero bytes behind the file for the bit reader
    size_t done = 0;
    while (done < size) {
        const ssize_t got = read(fd, image->data() + done, size - done);
        if (got < 0 && errno == EINTR) continue;
        if (got <= 0) break;
        done += (size_t)got;
    }
    close(fd);
    if (done != size) {
        set_error("%s: short read", path);
        return false;
    }
    return true;
}

int parse_stream(const char* path, const uint8_t* p, size_t n, StreamInfo* si) {
    size_t pos = 0;
    if (n >= 10 && memcmp(p, "ID3", 3) == 0) {            // ID3v2 tag in front of the stream
        const size_t tag = ((size_t)(p[6] & 0x7f) << 21) | ((size_t)(p[7] & 0x7f) << 14) |
                           ((size_t)(p[8] & 0x7f) << 7) | (size_t)(p[9] & 0x7f);
        pos = 10 + tag + ((p[5] & 0x10) ? 10 : 0);
    }
    if (pos + 4 > n || memcmp(p + pos, "fLaC", 4) != 0) {
        set_error("%s: not a FLAC stream", path);
        return PPGS_E_UNSUPPORTED;
    }
    pos += 4;
    bool first = true, last = false;
    while (!last) {
        if (pos + 4 > n) {
            set_error("%s: truncated FLAC metadata", path);
            return PPGS_E_INVALID;
        }
        last = (p[pos] & 0x80) != 0;
        const int type = p[pos] & 0x7f;
        const size_t length = ((size_t)p[pos + 1] << 16) | ((size_t)p[pos + 2] << 8) | p[pos + 3];
        pos += 4;
        if (pos + length > n || type == 127) {
            set_error("%s: bad FLAC metadata block", path);
            return PPGS_E_INVALID;
        }
        if (first) {
            if (type != 0 || length != 34) {
                set_error("%s: FLAC stream does not start with STREAMINFO", path);
                return PPGS_E_INVALID;
            }
            const uint8_t* s = p + pos;
            si->min_block = (s[0] << 8) | s[1];
            si->max_block = (s[2] << 8) | s[3];
            si->sample_rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
            si->channels = ((s[12] >> 1) & 7) + 1;
            si->bits = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
            si->total = ((int64_t)(s[13] & 15) << 32) | ((int64_t)s[14] << 24) | ((int64_t)s[15] << 16) |
                        ((int64_t)s[16] << 8) | (int64_t)s[17];
            memcpy(si->md5, s + 18, 16);
            for (int i = 0; i < 16; ++i) si->has_md5 |= si->md5[i] != 0;
            first = false;
        }
        pos += length;
    }
    si->audio_offset = pos;
    if (si->bits < 4 || si->sample_rate == 0) {
        set_error("%s: unusable STREAMINFO (%d bits, %d Hz)", path, si->bits, si->sample_rate);
        return PPGS_E_INVALID;
    }
    return PPGS_OK;
}

template <typename T>
const char* decode_residual(BitReader& br, int order, int block, T* out) {
    const int method = (int)br.read(2);
    if (method > 1) return "reserved residual coding method";
    const int param_bits = method == 0 ? 4 : 5;
    const uint32_t escape = method == 0 ? 15 : 31;
    const int partition_order = (int)br.read(4);
    const int partitions = 1 << partition_order;
    if (partition_order > 0 && (block & (partitions - 1))) return "block size not divisible into partitions";
    const int per_partition = block >> partition_order;
    if (per_partition < order) return "partition shorter than the predictor order";
    T* dst = out + order;
    for (int part = 0; part < partitions; ++part) {
        const int count = per_partition - (part == 0 ? order : 0);
        const uint32_t k = br.read(param_bits);
        if (k == escape) {
            const int raw_bits = (int)br.read(5);
            for (int i = 0; i < count; ++i) dst[i] = (T)br.read_signed(raw_bits);
        } else {
            for (int i = 0; i < count; ++i) {
                int64_t value;
                if (!br.read_rice((int)k, &value)) return "truncated residual";
                dst[i] = (T)value;
            }
        }
        dst += count;
        if (br.overrun()) return "truncated residual";
    }
    return nullptr;
}

// out[i] += (sum_j coefficient[j] * out[i - 1 - j]) >> shift, products and sum in 64 bits.
template <typename T, int ORDER>
void lpc_restore_fixed_order(T* out, int block, const int32_t* coefficient, int shift) {
    int64_t c[ORDER];
    for (int j = 0; j < ORDER; ++j) c[j] = coefficient[j];
    for (int i = ORDER; i < block; ++i) {
        int64_t sum = 0;
#pragma GCC unroll 32
        for (int j = ORDER - 1; j >= 0; --j) sum += c[j] * (int64_t)out[i - 1 - j];   // newest sample last
        out[i] = (T)(out[i] + (sum >> shift));
    }
}

template <typename T>
void lpc_restore(T* out, int block, int order, const int32_t* coefficient, int shift) {
    switch (order) {
#define PPGS_LPC_CASE(n) case n: lpc_restore_fixed_order<T, n>(out, block, coefficient, shift); return;
        PPGS_LPC_CASE(1) PPGS_LPC_CASE(2) PPGS_LPC_CASE(3) PPGS_LPC_CASE(4) PPGS_LPC_CASE(5) PPGS_LPC_CASE(6)
        PPGS_LPC_CASE(7) PPGS_LPC_CASE(8) PPGS_LPC_CASE(9) PPGS_LPC_CASE(10) PPGS_LPC_CASE(11) PPGS_LPC_CASE(12)
#undef PPGS_LPC_CASE
        default:
            for (int i = order; i < block; ++i) {
                int64_t sum = 0;
                for (int j = 0; j < order; ++j) sum += (int64_t)coefficient[j] * (int64_t)out[i - 1 - j];
                out[i] = (T)(out[i] + (sum >> shift));
            }
    }
}

// T = int32_t holds streams of <= 24 bits (25 with the side channel), int64_t the rest.
template <typename T>
const char* decode_subframe(BitReader& br, int bits, int block, T* out) {
    if (br.read(1)) return "subframe padding bit set";
    const int type = (int)br.read(6);
    int wasted = 0;
    if (br.read(1)) {
        uint32_t zeros;
        if (!br.read_unary(&zeros)) return "truncated subframe";
        wasted = (int)zeros + 1;
        if (wasted >= bits) return "wasted bits exceed the sample size";
        bits -= wasted;
    }
    if (type == 0) {
        const T value = (T)br.read_signed(bits);
        for (int i = 0; i < block; ++i) out[i] = value;
    } else if (type == 1) {
        for (int i = 0; i < block; ++i) out[i] = (T)br.read_signed(bits);
    } else if (type >= 8 && type <= 12) {
        const int order = type - 8;
        if (order > block) return "predictor order exceeds the block";
        for (int i = 0; i < order; ++i) out[i] = (T)br.read_signed(bits);
        if (const char* err = decode_residual(br, order, block, out)) return err;
        switch (order) {
            case 1:
                for (int i = 1; i < block; ++i) out[i] += out[i - 1];
                break;
            case 2:
                for (int i = 2; i < block; ++i) out[i] += 2 * out[i - 1] - out[i - 2];
                break;
            case 3:
                for (int i = 3; i < block; ++i) out[i] += 3 * out[i - 1] - 3 * out[i - 2] + out[i - 3];
                break;
            case 4:
                for (int i = 4; i < block; ++i)
                    out[i] += 4 * out[i - 1] - 6 * out[i - 2] + 4 * out[i - 3] - out[i - 4];
                break;
            default:
                break;
        }
    } else if (type >= 32) {
        const int order = (type & 31) + 1;
        if (order > block) return "predictor order exceeds the block";
        for (int i = 0; i < order; ++i) out[i] = (T)br.read_signed(bits);
        const int precision = (int)br.read(4) + 1;
        if (precision == 16) return "reserved predictor precision";
        const int shift = (int)br.read_signed(5);
        if (shift < 0) return "negative predictor shift";
        int32_t coefficient[32];
        for (int i = 0; i < order; ++i) coefficient[i] = (int32_t)br.read_signed(precision);
        if (const char* err = decode_residual(br, order, block, out)) return err;
        lpc_restore(out, block, order, coefficient, shift);
    } else {
        return "reserved subframe type";
    }
    if (wasted)
        for (int i = 0; i < block; ++i) out[i] = (T)(out[i] * ((T)1 << wasted));
    return br.overrun() ? "truncated subframe" : nullptr;
}

// Decodes every frame; dst (may be null: count / verify only) is channel-major with row stride
// `capacity`; dst16 (may be null; streams of <= 16 bits) takes channel 0 on the int16 scale, the
// form the file pipeline ships to the device.
template <typename T>
int decode_stream_as(const char* path, const uint8_t* p, size_t n, const StreamInfo& si, float* dst,
                     int16_t* dst16, int64_t capacity, int64_t* frames_out) {
    static const int kBlock[16] = {0, 192, 576, 1152, 2304, 4608, 0, 0, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768};
    static const int kRate[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
    static const int kBits[8] = {0, 8, 12, -1, 16, 20, 24, 32};
    std::vector<T> channel[8];
    std::vector<uint8_t> pcm;
    Md5 md5;
    const int sample_bytes = (si.bits + 7) / 8;
    const float scale = 1.0f / (float)((int64_t)1 << (si.bits - 1));
    int64_t done = 0;
    size_t pos = si.audio_offset;
    while (pos < n && !(si.total > 0 && done >= si.total)) {
        if (n - pos == 128 && memcmp(p + pos, "TAG", 3) == 0) break;          // ID3v1 trailer
        if (pos + 6 > n || p[pos] != 0xff || (p[pos + 1] & 0xfe) != 0xf8) {
            set_error("%s: lost FLAC frame sync at byte %zu", path, pos);
            return PPGS_E_INVALID;
        }
        const int block_code = p[pos + 2] >> 4, rate_code = p[pos + 2] & 15;
        const int assignment = p[pos + 3] >> 4, size_code = (p[pos + 3] >> 1) & 7;
        if ((p[pos + 3] & 1) || block_code == 0 || rate_code == 15 || assignment > 10 || kBits[size_code] < 0) {
            set_error("%s: reserved value in the frame header at byte %zu", path, pos);
            return PPGS_E_INVALID;
        }
        size_t cur = pos + 4;
        int block = kBlock[block_code];
        // frame / sample number: UTF-8-style, 1..7 bytes
        if (cur >= n) goto truncated;
        {
            int ones = 0;
            while (ones < 8 && (p[cur] & (0x80 >> ones))) ++ones;
            if (ones == 1 || ones == 8) {
                set_error("%s: bad coded number in the frame header at byte %zu", path, pos);
                return PPGS_E_INVALID;
            }
            cur += ones == 0 ? 1 : (size_t)ones;
        }
        if (block_code == 6) {
            if (cur + 1 > n) goto truncated;
            block = p[cur] + 1;
            cur += 1;
        } else if (block_code == 7) {
            if (cur + 2 > n) goto truncated;
            block = ((p[cur] << 8) | p[cur + 1]) + 1;
            cur += 2;
        }
        {
            int rate = rate_code < 12 ? kRate[rate_code] : 0;
            if (rate_code == 12) {
                if (cur + 1 > n) goto truncated;
                rate = p[cur] * 1000;
                cur += 1;
            } else if (rate_code == 13 || rate_code == 14) {
                if (cur + 2 > n) goto truncated;
                rate = ((p[cur] << 8) | p[cur + 1]) * (rate_code == 14 ? 10 : 1);
                cur += 2;
            }
            if (cur + 1 > n) goto truncated;
            if (crc8(p + pos, cur - pos) != p[cur]) {
                set_error("%s: frame header CRC mismatch at byte %zu", path, pos);
                return PPGS_E_INVALID;
            }
            cur += 1;
            const int channels = assignment < 8 ? assignment + 1 : 2;
            const int bits = size_code == 0 ? si.bits : kBits[size_code];
            if (channels != si.channels || bits != si.bits || (rate != 0 && rate != si.sample_rate)) {
                set_error("%s: frame at byte %zu changes the stream format (%d ch, %d bits, %d Hz)", path,
                          pos, channels, bits, rate);
                return PPGS_E_UNSUPPORTED;
            }
            BitReader br(p, n, cur);
            for (int c = 0; c < channels; ++c) {
                if ((int)channel[c].size() < block) channel[c].resize((size_t)block);
                const bool side = (assignment == 8 && c == 1) || (assignment == 9 && c == 0) ||
                                  (assignment == 10 && c == 1);
                if (const char* err = decode_subframe(br, bits + (side ? 1 : 0), block, channel[c].data())) {
                    set_error("%s: %s (frame at byte %zu, channel %d)", path, err, pos, c);
                    return PPGS_E_INVALID;
                }
            }
            br.align();
            const size_t end = br.byte_position();
            if (end + 2 > n) goto truncated;
            if (crc16(p + pos, end - pos) != (uint16_t)((p[end] << 8) | p[end + 1])) {
                set_error("%s: frame CRC mismatch (frame at byte %zu)", path, pos);
                return PPGS_E_INVALID;
            }
            T* a = channel[0].data();
            T* b = channels > 1 ? channel[1].data() : nullptr;
            if (assignment == 8) {
                for (int i = 0; i < block; ++i) b[i] = a[i] - b[i];
            } else if (assignment == 9) {
                for (int i = 0; i < block; ++i) a[i] += b[i];
            } else if (assignment == 10) {
                for (int i = 0; i < block; ++i) {
                    const int64_t side = b[i];
                    const int64_t mid = (int64_t)a[i] * 2 + (side & 1);
                    a[i] = (T)((mid + side) >> 1);
                    b[i] = (T)((mid - side) >> 1);
                }
            }
            int keep = block;
            if (si.total > 0 && done + keep > si.total) keep = (int)(si.total - done);
            if (dst16) {
                if (done + keep > capacity) {
                    set_error("%s: more than %lld frames; the buffer is too small", path, (long long)capacity);
                    return PPGS_E_INVALID;
                }
                const int up = 16 - si.bits;
                const T* src = channel[0].data();
                for (int i = 0; i < keep; ++i) dst16[done + i] = (int16_t)(src[i] * ((T)1 << up));
            }
            if (dst) {
                if (done + keep > capacity) {
                    set_error("%s: more than %lld frames; the buffer is too small", path, (long long)capacity);
                    return PPGS_E_INVALID;
                }
                for (int c = 0; c < channels; ++c) {
                    float* row = dst + (int64_t)c * capacity + done;
                    const T* src = channel[c].data();
                    for (int i = 0; i < keep; ++i) row[i] = (float)src[i] * scale;
                }
            }
            if (si.has_md5) {
                pcm.resize((size_t)block * channels * sample_bytes);
                uint8_t* w = pcm.data();
                if (channels == 1 && sample_bytes == 2) {
                    const T* src = channel[0].data();
                    for (int i = 0; i < block; ++i) {
                        const int16_t v = (int16_t)src[i];
                        memcpy(w + 2 * i, &v, 2);       // little-endian host
                    }
                } else {
                    for (int i = 0; i < block; ++i)
                        for (int c = 0; c < channels; ++c) {
                            const int64_t v = channel[c][(size_t)i];
                            for (int byte = 0; byte < sample_bytes; ++byte) *w++ = (uint8_t)(v >> (8 * byte));
                        }
                }
                md5.update(pcm.data(), pcm.size());
            }
            done += keep;
            pos = end + 2;
        }
        continue;
    truncated:
        set_error("%s: truncated FLAC frame at byte %zu", path, pos);
        return PPGS_E_INVALID;
    }
    if (si.total > 0 && done != si.total) {
        set_error("%s: decoded %lld of %lld frames", path, (long long)done, (long long)si.total);
        return PPGS_E_INVALID;
    }
    if (si.has_md5) {
        uint8_t digest[16];
        md5.finish(digest);
        if (memcmp(digest, si.md5, 16) != 0) {
            set_error("%s: MD5 of the decoded signal does not match STREAMINFO", path);
            return PPGS_E_INVALID;
        }
    }
    if (frames_out) *frames_out = done;
    return PPGS_OK;
}

int decode_stream(const char* path, const uint8_t* p, size_t n, const StreamInfo& si, float* dst,
                  int16_t* dst16, int64_t capacity, int64_t* frames_out) {
    if (si.bits <= 24) return decode_stream_as<int32_t>(path, p, n, si, dst, dst16, capacity, frames_out);
    return decode_stream_as<int64_t>(path, p, n, si, dst, dst16, capacity, frames_out);
}

// STREAMINFO from the first bytes of the file only (header probes of a corpus must not read it).
int probe_streaminfo(const char* path, StreamInfo* si) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) {
        set_error("%s: cannot open: %s", path, strerror(errno));
        return PPGS_E_INVALID;
    }
    uint8_t head[10 + 42];
    ssize_t got = pread(fd, head, 10, 0);
    off_t at = 0;
    if (got == 10 && memcmp(head, "ID3", 3) == 0)
        at = 10 + (((off_t)(head[6] & 0x7f) << 21) | ((off_t)(head[7] & 0x7f) << 14) |
                   ((off_t)(head[8] & 0x7f) << 7) | (off_t)(head[9] & 0x7f)) + ((head[5] & 0x10) ? 10 : 0);
    got = pread(fd, head, 42, at);
    close(fd);
    if (got < 4 || memcmp(head, "fLaC", 4) != 0) {
        set_error("%s: not a FLAC stream", path);
        return PPGS_E_UNSUPPORTED;
    }
    if (got < 42) {
        set_error("%s: truncated FLAC metadata", path);
        return PPGS_E_INVALID;
    }
    // a complete one-block stream for parse_stream: mark STREAMINFO as the last block
    head[4] |= 0x80;
    return parse_stream(path, head, 42, si);
}

}  // namespace

// File pipeline (io.cu): channel 0 of a 16 kHz FLAC file of <= 16 bits as int16, zero padded to
// `capacity`; the stream must hold exactly `expect_frames` frames.
int flac_read_pcm16(const char* path, int16_t* dst, int64_t capacity, int64_t expect_frames) {
    std::vector<uint8_t> image;
    if (!load_file(path, &image)) return PPGS_E_INVALID;
    StreamInfo si;
    PPGS_CHECK(parse_stream(path, image.data(), image.size() - kImagePad, &si));
    if (si.bits > 16 || si.sample_rate != 16000) {
        set_error("%s: the native file pipeline takes FLAC of <= 16 bits at 16 kHz (got %d bits, %d Hz)", path,
                  si.bits, si.sample_rate);
        return PPGS_E_UNSUPPORTED;
    }
    int64_t total = 0;
    PPGS_CHECK(decode_stream(path, image.data(), image.size() - kImagePad, si, nullptr, dst, capacity, &total));
    if (total != expect_frames) {
        set_error("%s: %lld frames, expected %lld", path, (long long)total, (long long)expect_frames);
        return PPGS_E_UNSUPPORTED;
    }
    if (total < capacity) memset(dst + total, 0, (size_t)(capacity - total) * sizeof(int16_t));
    return PPGS_OK;
}

}  // namespace ppgs

using namespace ppgs;

extern "C" {

int ppgs_flac_info(const char* path, int64_t* frames, int* sample_rate, int* channels, int* bits) {
    if (!path) {
        set_error("flac_info: bad argument");
        return PPGS_E_INVALID;
    }
    StreamInfo si;
    std::vector<uint8_t> image;
    int rc = probe_streaminfo(path, &si);
    if (rc != PPGS_OK) return rc;
    int64_t total = si.total;
    if (total == 0) {    // length not recorded (streamed encoder): count by decoding
        if (!load_file(path, &image)) return PPGS_E_INVALID;
        PPGS_CHECK(parse_stream(path, image.data(), image.size() - kImagePad, &si));
        PPGS_CHECK(decode_stream(path, image.data(), image.size() - kImagePad, si, nullptr, nullptr, 0, &total));
    }
    if (frames) *frames = total;
    if (sample_rate) *sample_rate = si.sample_rate;
    if (channels) *channels = si.channels;
    if (bits) *bits = si.bits;
    return PPGS_OK;
}

int ppgs_flac_read_f32(const char* path, float* dst, int64_t capacity, int64_t* frames, int* sample_rate,
                       int* channels) {
    if (!path || !dst || capacity < 0) {
        set_error("flac_read_f32: bad argument");
        return PPGS_E_INVALID;
    }
    std::vector<uint8_t> image;
    if (!load_file(path, &image)) return PPGS_E_INVALID;
    StreamInfo si;
    PPGS_CHECK(parse_stream(path, image.data(), image.size() - kImagePad, &si));
    int64_t total = 0;
    PPGS_CHECK(decode_stream(path, image.data(), image.size() - kImagePad, si, dst, nullptr, capacity, &total));
    if (frames) *frames = total;
    if (sample_rate) *sample_rate = si.sample_rate;
    if (channels) *channels = si.channels;
    return PPGS_OK;
}

}  // extern "C"
