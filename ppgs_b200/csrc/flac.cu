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
        static const uint32_t K[64] = {
            0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
            0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
            0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
            0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
            0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
            0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
            0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
            0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
        static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                                  5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
        uint32_t m[16];
        for (int i = 0; i < 16; ++i)
            m[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
                   ((uint32_t)p[4 * i + 3] << 24);
        uint32_t A = a, B = b, C = c, D = d;
        for (int i = 0; i < 64; ++i) {
            uint32_t f;
            int g;
            if (i < 16) {
                f = (B & C) | (~B & D);
                g = i;
            } else if (i < 32) {
                f = (D & B) | (~D & C);
                g = (5 * i + 1) & 15;
            } else if (i < 48) {
                f = B ^ C ^ D;
                g = (3 * i + 5) & 15;
            } else {
                f = C ^ (B | ~D);
                g = (7 * i) & 15;
            }
            const uint32_t next = B + rotl(A + f + K[i] + m[g], S[i]);
            A = D;
            D = C;
            C = B;
            B = next;
        }
        a += A;
        b += B;
        c += C;
        d += D;
    }

    void update(const uint8_t* p, size_t n) {
        bytes += n;
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
    uint16_t crc16[256];
    CrcTables() {
        for (int i = 0; i < 256; ++i) {
            uint8_t c8 = (uint8_t)i;
            uint16_t c16 = (uint16_t)(i << 8);
            for (int bit = 0; bit < 8; ++bit) {
                c8 = (uint8_t)((c8 << 1) ^ ((c8 & 0x80) ? 0x07 : 0));               // x^8 + x^2 + x + 1
                c16 = (uint16_t)((c16 << 1) ^ ((c16 & 0x8000) ? 0x8005 : 0));       // x^16 + x^15 + x^2 + 1
            }
            crc8[i] = c8;
            crc16[i] = c16;
        }
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
    for (size_t i = 0; i < n; ++i) c = (uint16_t)((c << 8) ^ t.crc16[(c >> 8) ^ p[i]]);
    return c;
}

// ---- MSB-first bit reader over the file image ------------------------------------------------
struct BitReader {
    const uint8_t* data;
    size_t size;
    size_t next;         // next byte to load
    uint64_t window = 0; // valid bits are the top `bits`
    int bits = 0;

    BitReader(const uint8_t* d, size_t n, size_t start) : data(d), size(n), next(start) {}

    void refill() {
        while (bits <= 56) {
            const uint64_t byte = next < size ? data[next] : 0;
            ++next;
            window |= byte << (56 - bits);
            bits += 8;
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
    image->resize((size_t)st.st_size);
    size_t done = 0;
    while (done < image->size()) {
        const ssize_t got = read(fd, image->data() + done, image->size() - done);
        if (got < 0 && errno == EINTR) continue;
        if (got <= 0) break;
        done += (size_t)got;
    }
    close(fd);
    if (done != image->size()) {
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

const char* decode_residual(BitReader& br, int order, int block, int64_t* out) {
    const int method = (int)br.read(2);
    if (method > 1) return "reserved residual coding method";
    const int param_bits = method == 0 ? 4 : 5;
    const uint32_t escape = method == 0 ? 15 : 31;
    const int partition_order = (int)br.read(4);
    const int partitions = 1 << partition_order;
    if (partition_order > 0 && (block & (partitions - 1))) return "block size not divisible into partitions";
    const int per_partition = block >> partition_order;
    if (per_partition < order) return "partition shorter than the predictor order";
    int64_t* dst = out + order;
    for (int part = 0; part < partitions; ++part) {
        const int count = per_partition - (part == 0 ? order : 0);
        const uint32_t k = br.read(param_bits);
        if (k == escape) {
            const int raw_bits = (int)br.read(5);
            for (int i = 0; i < count; ++i) dst[i] = br.read_signed(raw_bits);
        } else {
            for (int i = 0; i < count; ++i) {
                uint32_t quotient;
                if (!br.read_unary(&quotient)) return "truncated residual";
                const uint64_t folded = ((uint64_t)quotient << k) | br.read((int)k);
                dst[i] = (int64_t)(folded >> 1) ^ -(int64_t)(folded & 1);
            }
        }
        dst += count;
        if (br.overrun()) return "truncated residual";
    }
    return nullptr;
}

const char* decode_subframe(BitReader& br, int bits, int block, int64_t* out) {
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
        const int64_t value = br.read_signed(bits);
        for (int i = 0; i < block; ++i) out[i] = value;
    } else if (type == 1) {
        for (int i = 0; i < block; ++i) out[i] = br.read_signed(bits);
    } else if (type >= 8 && type <= 12) {
        const int order = type - 8;
        if (order > block) return "predictor order exceeds the block";
        for (int i = 0; i < order; ++i) out[i] = br.read_signed(bits);
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
        for (int i = 0; i < order; ++i) out[i] = br.read_signed(bits);
        const int precision = (int)br.read(4) + 1;
        if (precision == 16) return "reserved predictor precision";
        const int shift = (int)br.read_signed(5);
        if (shift < 0) return "negative predictor shift";
        int64_t coefficient[32];
        for (int i = 0; i < order; ++i) coefficient[i] = br.read_signed(precision);
        if (const char* err = decode_residual(br, order, block, out)) return err;
        for (int i = order; i < block; ++i) {
            int64_t sum = 0;
            for (int j = 0; j < order; ++j) sum += coefficient[j] * out[i - 1 - j];
            out[i] += sum >> shift;
        }
    } else {
        return "reserved subframe type";
    }
    if (wasted)
        for (int i = 0; i < block; ++i) out[i] *= (int64_t)1 << wasted;
    return br.overrun() ? "truncated subframe" : nullptr;
}

// Decodes every frame; dst (may be null: count / verify only) is channel-major with row stride
// `capacity`.
int decode_stream(const char* path, const uint8_t* p, size_t n, const StreamInfo& si, float* dst,
                  int64_t capacity, int64_t* frames_out) {
    static const int kBlock[16] = {0, 192, 576, 1152, 2304, 4608, 0, 0, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768};
    static const int kRate[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
    static const int kBits[8] = {0, 8, 12, -1, 16, 20, 24, 32};
    std::vector<int64_t> channel[8];
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
            int64_t* a = channel[0].data();
            int64_t* b = channels > 1 ? channel[1].data() : nullptr;
            if (assignment == 8) {
                for (int i = 0; i < block; ++i) b[i] = a[i] - b[i];
            } else if (assignment == 9) {
                for (int i = 0; i < block; ++i) a[i] += b[i];
            } else if (assignment == 10) {
                for (int i = 0; i < block; ++i) {
                    const int64_t side = b[i];
                    const int64_t mid = a[i] * 2 + (side & 1);
                    a[i] = (mid + side) >> 1;
                    b[i] = (mid - side) >> 1;
                }
            }
            int keep = block;
            if (si.total > 0 && done + keep > si.total) keep = (int)(si.total - done);
            if (dst) {
                if (done + keep > capacity) {
                    set_error("%s: more than %lld frames; the buffer is too small", path, (long long)capacity);
                    return PPGS_E_INVALID;
                }
                for (int c = 0; c < channels; ++c) {
                    float* row = dst + (int64_t)c * capacity + done;
                    const int64_t* src = channel[c].data();
                    for (int i = 0; i < keep; ++i) row[i] = (float)src[i] * scale;
                }
            }
            if (si.has_md5) {
                pcm.resize((size_t)block * channels * sample_bytes);
                uint8_t* w = pcm.data();
                for (int i = 0; i < block; ++i)
                    for (int c = 0; c < channels; ++c) {
                        const int64_t v = channel[c][(size_t)i];
                        for (int byte = 0; byte < sample_bytes; ++byte) *w++ = (uint8_t)(v >> (8 * byte));
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

}  // namespace
}  // namespace ppgs

using namespace ppgs;

extern "C" {

int ppgs_flac_info(const char* path, int64_t* frames, int* sample_rate, int* channels, int* bits) {
    if (!path) {
        set_error("flac_info: bad argument");
        return PPGS_E_INVALID;
    }
    std::vector<uint8_t> image;
    if (!load_file(path, &image)) return PPGS_E_INVALID;
    StreamInfo si;
    PPGS_CHECK(parse_stream(path, image.data(), image.size(), &si));
    int64_t total = si.total;
    if (total == 0)      // length not recorded (streamed encoder): count by decoding
        PPGS_CHECK(decode_stream(path, image.data(), image.size(), si, nullptr, 0, &total));
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
    PPGS_CHECK(parse_stream(path, image.data(), image.size(), &si));
    int64_t total = 0;
    PPGS_CHECK(decode_stream(path, image.data(), image.size(), si, dst, capacity, &total));
    if (frames) *frames = total;
    if (sample_rate) *sample_rate = si.sample_rate;
    if (channels) *channels = si.channels;
    return PPGS_OK;
}

}  // extern "C"
