// Audio ingest and posteriorgram egress of the file API:
//
//   * RIFF/WAVE header probe and PCM decode      (torchaudio.info / torchaudio.load as
//     called from ppgs/data/dataset.py:187 and ppgs/load.py:17-30)
//   * sinc-interpolation resampler on the GPU    (torchaudio.transforms.Resample as called
//     from ppgs/core.py:599-608; algorithm restated from torchaudio 2.11
//     functional._get_sinc_resample_kernel / _apply_sinc_resample_kernel)
//   * int16 PCM -> fp32 on the GPU               (the /32768 normalisation of torchaudio.load)
//   * torch.load-compatible `.pt` writer         (preprocess.save_masked = torch.save of the
//     cropped tensor, ppgs/preprocess/core.py:219-221)
//   * ppgs_files_to_files: the batching loop of ppgs/core.py:280-391 as a native pipeline:
//     reader threads -> pinned int16 batches -> H2D -> mel + Transformer -> D2H -> writer
//     threads, all inside one C-ABI call, no Python in the loop.
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <string.h>
#include <strings.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace ppgs {

// ---------------------------------------------------------------------------
// RIFF / WAVE
// ---------------------------------------------------------------------------
struct WavHeader {
    int64_t frames = 0;        // samples per channel
    int sample_rate = 0;
    int channels = 0;
    int bits = 0;
    int is_float = 0;
    int64_t data_offset = 0;   // byte offset of the first sample
    int64_t data_bytes = 0;
};

static uint32_t rd32(const unsigned char* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static uint32_t rd16(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

static bool pread_all(int fd, void* dst, size_t bytes, int64_t offset) {
    char* out = static_cast<char*>(dst);
    while (bytes) {
        ssize_t got = pread(fd, out, bytes, offset);
        if (got < 0 && errno == EINTR) continue;
        if (got <= 0) return false;
        out += got;
        offset += got;
        bytes -= (size_t)got;
    }
    return true;
}

// Walks the chunk list; tolerates LIST / fact / bext chunks before `data`, and a data size
// of 0 or 0xFFFFFFFF (streamed writers) by clamping to the file size.
static int parse_wav(int fd, const char* path, WavHeader* h) {
    struct stat st;
    if (fstat(fd, &st) != 0) {
        set_error("%s: fstat failed: %s", path, strerror(errno));
        return PPGS_E_INVALID;
    }
    unsigned char head[12];
    if (!pread_all(fd, head, 12, 0) || memcmp(head, "RIFF", 4) || memcmp(head + 8, "WAVE", 4)) {
        set_error("%s: not a RIFF/WAVE file", path);
        return PPGS_E_UNSUPPORTED;
    }
    int64_t at = 12;
    bool have_fmt = false;
    while (at + 8 <= st.st_size) {
        unsigned char ck[8];
        if (!pread_all(fd, ck, 8, at)) break;
        const int64_t size = rd32(ck + 4);
        if (!memcmp(ck, "fmt ", 4)) {
            unsigned char fmt[40] = {0};
            const size_t take = (size_t)(size < 40 ? size : 40);
            if (size < 16 || !pread_all(fd, fmt, take, at + 8)) {
                set_error("%s: truncated fmt chunk", path);
                return PPGS_E_INVALID;
            }
            uint32_t tag = rd16(fmt);
            h->channels = (int)rd16(fmt + 2);
            h->sample_rate = (int)rd32(fmt + 4);
            h->bits = (int)rd16(fmt + 14);
            if (tag == 0xFFFE && size >= 26) tag = rd16(fmt + 24);   // WAVE_FORMAT_EXTENSIBLE
            if (tag != 1 && tag != 3) {
                set_error("%s: WAVE format tag %u is not PCM / IEEE float", path, tag);
                return PPGS_E_UNSUPPORTED;
            }
            h->is_float = tag == 3;
            have_fmt = true;
        } else if (!memcmp(ck, "data", 4)) {
            if (!have_fmt) {
                set_error("%s: data chunk before fmt chunk", path);
                return PPGS_E_INVALID;
            }
            h->data_offset = at + 8;
            int64_t bytes = size;
            if (bytes == 0 || bytes == 0xFFFFFFFFll || h->data_offset + bytes > st.st_size)
                bytes = st.st_size - h->data_offset;
            const int frame_bytes = h->channels * (h->bits / 8);
            const bool pcm_ok = !h->is_float && (h->bits == 8 || h->bits == 16 || h->bits == 24 || h->bits == 32);
            const bool float_ok = h->is_float && (h->bits == 32 || h->bits == 64);
            if (h->channels <= 0 || frame_bytes <= 0 || !(pcm_ok || float_ok)) {
                set_error("%s: unsupported sample layout (%d channels, %d bits%s)", path, h->channels,
                          h->bits, h->is_float ? " float" : "");
                return PPGS_E_UNSUPPORTED;
            }
            h->frames = bytes / frame_bytes;
            h->data_bytes = h->frames * frame_bytes;
            return PPGS_OK;
        }
        at += 8 + size + (size & 1);
    }
    set_error("%s: no data chunk", path);
    return PPGS_E_INVALID;
}

struct Fd {
    int fd;
    explicit Fd(const char* path, int flags, int mode = 0) : fd(open(path, flags, mode)) {}
    ~Fd() {
        if (fd >= 0) close(fd);
    }
};

// channel 0 of `frames` frames starting at frame 0, normalised like torchaudio.load
static int decode_channel0(const WavHeader& h, const unsigned char* raw, int64_t frames, float* dst) {
    const int step = h.channels * (h.bits / 8);
    if (h.is_float && h.bits == 32) {
        for (int64_t i = 0; i < frames; ++i) memcpy(dst + i, raw + i * step, 4);
    } else if (h.is_float) {
        for (int64_t i = 0; i < frames; ++i) {
            double v;
            memcpy(&v, raw + i * step, 8);
            dst[i] = (float)v;
        }
    } else if (h.bits == 16) {
        for (int64_t i = 0; i < frames; ++i) {
            int16_t v;
            memcpy(&v, raw + i * step, 2);
            dst[i] = (float)v * (1.0f / 32768.0f);
        }
    } else if (h.bits == 8) {
        for (int64_t i = 0; i < frames; ++i) dst[i] = ((float)raw[i * step] - 128.0f) * (1.0f / 128.0f);
    } else if (h.bits == 24) {
        for (int64_t i = 0; i < frames; ++i) {
            const unsigned char* p = raw + i * step;
            const int32_t v = (int32_t)(((uint32_t)p[0] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 24));
            dst[i] = (float)v * (1.0f / 2147483648.0f);
        }
    } else {
        for (int64_t i = 0; i < frames; ++i) {
            int32_t v;
            memcpy(&v, raw + i * step, 4);
            dst[i] = (float)v * (1.0f / 2147483648.0f);
        }
    }
    return PPGS_OK;
}

// channel 0 of a 16-bit PCM file straight into `dst` (no conversion)
static int read_pcm16_channel0(int fd, const char* path, const WavHeader& h, int64_t frames, int16_t* dst,
                               std::vector<int16_t>& scratch) {
    if (h.channels == 1) {
        if (!pread_all(fd, dst, (size_t)frames * 2, h.data_offset)) {
            set_error("%s: short read", path);
            return PPGS_E_INVALID;
        }
        return PPGS_OK;
    }
    scratch.resize((size_t)frames * h.channels);
    if (!pread_all(fd, scratch.data(), scratch.size() * 2, h.data_offset)) {
        set_error("%s: short read", path);
        return PPGS_E_INVALID;
    }
    for (int64_t i = 0; i < frames; ++i) dst[i] = scratch[(size_t)i * h.channels];
    return PPGS_OK;
}

// ---------------------------------------------------------------------------
// torch.save-compatible writer: a stored (uncompressed) zip archive
//   <stem>/data.pkl   pickle protocol 2: torch._utils._rebuild_tensor_v2(FloatStorage '0', ...)
//   <stem>/byteorder  "little"
//   <stem>/data/0     the fp32 storage, 64-byte aligned in the file like torch's writer
//   <stem>/version    "3\n"
// ---------------------------------------------------------------------------
static uint32_t crc_table[8][256];
static std::once_flag crc_once;

static void crc_init() {
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        crc_table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
        for (int t = 1; t < 8; ++t)
            crc_table[t][i] = (crc_table[t - 1][i] >> 8) ^ crc_table[0][crc_table[t - 1][i] & 0xFF];
}

static uint32_t crc32_bytes(const void* data, size_t bytes) {
    std::call_once(crc_once, crc_init);
    const unsigned char* p = static_cast<const unsigned char*>(data);
    uint32_t c = 0xFFFFFFFFu;
    while (bytes >= 8) {
        uint32_t a, b;
        memcpy(&a, p, 4);
        memcpy(&b, p + 4, 4);
        a ^= c;
        c = crc_table[7][a & 0xFF] ^ crc_table[6][(a >> 8) & 0xFF] ^ crc_table[5][(a >> 16) & 0xFF] ^
            crc_table[4][a >> 24] ^ crc_table[3][b & 0xFF] ^ crc_table[2][(b >> 8) & 0xFF] ^
            crc_table[1][(b >> 16) & 0xFF] ^ crc_table[0][b >> 24];
        p += 8;
        bytes -= 8;
    }
    while (bytes--) c = crc_table[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

struct ByteSink {
    std::vector<unsigned char> b;
    void u8(unsigned v) { b.push_back((unsigned char)v); }
    void u16(unsigned v) {
        u8(v & 0xFF);
        u8((v >> 8) & 0xFF);
    }
    void u32(uint32_t v) {
        u16(v & 0xFFFF);
        u16(v >> 16);
    }
    void raw(const void* p, size_t n) {
        const unsigned char* s = static_cast<const unsigned char*>(p);
        b.insert(b.end(), s, s + n);
    }
    void str(const char* s) { raw(s, strlen(s)); }
};

static void pickle_int(ByteSink& s, int64_t v) {
    if (v >= 0 && v < 256) {
        s.u8('K');
        s.u8((unsigned)v);
    } else if (v >= 0 && v < 65536) {
        s.u8('M');
        s.u16((unsigned)v);
    } else {
        s.u8('J');
        s.u32((uint32_t)(int32_t)v);
    }
}

static void pickle_unicode(ByteSink& s, const char* text) {
    s.u8('X');
    s.u32((uint32_t)strlen(text));
    s.str(text);
}

// The pickle torch.save emits for one contiguous fp32 CPU tensor of shape (rows, cols)
// (memo PUTs omitted; the unpickler does not need them).
static void tensor_pickle(ByteSink& s, int64_t rows, int64_t cols, const char* storage_global) {
    s.u8(0x80);
    s.u8(2);
    s.str("ctorch._utils\n_rebuild_tensor_v2\n");
    s.u8('(');
    s.u8('(');
    pickle_unicode(s, "storage");
    s.str(storage_global);
    pickle_unicode(s, "0");
    pickle_unicode(s, "cpu");
    pickle_int(s, rows * cols);
    s.u8('t');
    s.u8('Q');
    pickle_int(s, 0);
    pickle_int(s, rows);
    pickle_int(s, cols);
    s.u8(0x86);
    pickle_int(s, cols);
    pickle_int(s, 1);
    s.u8(0x86);
    s.u8(0x89);
    s.str("ccollections\nOrderedDict\n");
    s.u8(')');
    s.u8('R');
    s.u8('t');
    s.u8('R');
    s.u8('.');
}

struct ZipEntry {
    std::string name;
    uint32_t crc, size, offset;
};

static void zip_local(ByteSink& out, std::vector<ZipEntry>& dir, const std::string& name, const void* data,
                      size_t bytes, uint32_t crc, bool align64, bool append_data) {
    ZipEntry entry{name, crc, (uint32_t)bytes, (uint32_t)out.b.size()};
    size_t extra = 0;
    if (align64) {   // pad with an "FB" extra field so that the payload starts 64-byte aligned
        const size_t start = out.b.size() + 30 + name.size() + 4;
        extra = 4 + ((64 - start % 64) % 64);
    }
    out.u32(0x04034b50);
    out.u16(20);
    out.u16(0x0800);   // UTF-8 names
    out.u16(0);        // stored
    out.u16(0);
    out.u16(0x21);     // 1980-01-01
    out.u32(crc);
    out.u32((uint32_t)bytes);
    out.u32((uint32_t)bytes);
    out.u16((unsigned)name.size());
    out.u16((unsigned)extra);
    out.str(name.c_str());
    if (extra) {
        out.u8('F');
        out.u8('B');
        out.u16((unsigned)(extra - 4));
        for (size_t i = 4; i < extra; ++i) out.u8('Z');
    }
    if (append_data) out.raw(data, bytes);
    dir.push_back(entry);
}

static void zip_central(ByteSink& out, const std::vector<ZipEntry>& dir, uint32_t dir_offset) {
    ByteSink cd;
    for (const ZipEntry& e : dir) {
        cd.u32(0x02014b50);
        cd.u16(20);
        cd.u16(20);
        cd.u16(0x0800);
        cd.u16(0);
        cd.u16(0);
        cd.u16(0x21);
        cd.u32(e.crc);
        cd.u32(e.size);
        cd.u32(e.size);
        cd.u16((unsigned)e.name.size());
        cd.u16(0);
        cd.u16(0);
        cd.u16(0);
        cd.u16(0);
        cd.u32(0);
        cd.u32(e.offset);
        cd.str(e.name.c_str());
    }
    out.raw(cd.b.data(), cd.b.size());
    out.u32(0x06054b50);
    out.u16(0);
    out.u16(0);
    out.u16((unsigned)dir.size());
    out.u16((unsigned)dir.size());
    out.u32((uint32_t)cd.b.size());
    out.u32(dir_offset);
    out.u16(0);
}

static bool write_all(int fd, const void* data, size_t bytes) {
    const char* p = static_cast<const char*>(data);
    while (bytes) {
        ssize_t put = write(fd, p, bytes);
        if (put < 0 && errno == EINTR) continue;
        if (put <= 0) return false;
        p += put;
        bytes -= (size_t)put;
    }
    return true;
}

static std::string archive_stem(const char* path) {
    std::string s(path);
    const size_t slash = s.find_last_of('/');
    if (slash != std::string::npos) s = s.substr(slash + 1);
    const size_t dot = s.find_last_of('.');
    if (dot != std::string::npos && dot > 0) s = s.substr(0, dot);
    return s.empty() ? std::string("archive") : s;
}

// `data`: rows x cols fp32, row stride `row_stride` elements (cropping a padded batch row
// is a strided read here, the file always holds the contiguous tensor).
// `elem` = bytes per element (4: torch.FloatStorage, 2: torch.HalfStorage)
static int pt_write(const char* path, const void* data_, int64_t rows, int64_t cols, int64_t row_stride,
                    std::vector<float>& contiguous, ByteSink& head, int elem = 4) {
    const char* data = static_cast<const char*>(data_);
    if (rows < 0 || cols < 0 || row_stride < cols || rows * cols > (int64_t)0x3FFFFFFF) {
        set_error("pt_write: bad shape (%lld, %lld)", (long long)rows, (long long)cols);
        return PPGS_E_INVALID;
    }
    const char* payload = data;
    if (row_stride != cols) {
        contiguous.resize((size_t)(rows * cols * elem + 3) / 4);
        char* packed = reinterpret_cast<char*>(contiguous.data());
        for (int64_t r = 0; r < rows; ++r)
            memcpy(packed + r * cols * elem, data + r * row_stride * elem, (size_t)cols * elem);
        payload = packed;
    }
    const size_t payload_bytes = (size_t)(rows * cols) * elem;
    const std::string stem = archive_stem(path);
    head.b.clear();
    std::vector<ZipEntry> dir;
    ByteSink pkl;
    tensor_pickle(pkl, rows, cols, elem == 2 ? "ctorch\nHalfStorage\n" : "ctorch\nFloatStorage\n");
    zip_local(head, dir, stem + "/data.pkl", pkl.b.data(), pkl.b.size(), crc32_bytes(pkl.b.data(), pkl.b.size()),
              false, true);
    zip_local(head, dir, stem + "/byteorder", "little", 6, crc32_bytes("little", 6), false, true);
    zip_local(head, dir, stem + "/data/0", payload, payload_bytes, crc32_bytes(payload, payload_bytes), true,
              false);
    ByteSink tail;
    const uint32_t after_payload = (uint32_t)(head.b.size() + payload_bytes);
    {   // entries after the payload are assembled with offsets relative to the whole file
        ByteSink scratch;
        scratch.b.resize(after_payload);   // placeholder so that offsets come out right
        zip_local(scratch, dir, stem + "/version", "3\n", 2, crc32_bytes("3\n", 2), false, true);
        const uint32_t dir_offset = (uint32_t)scratch.b.size();
        zip_central(scratch, dir, dir_offset);
        tail.raw(scratch.b.data() + after_payload, scratch.b.size() - after_payload);
    }
    Fd out(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (out.fd < 0) {
        set_error("%s: cannot open for writing: %s", path, strerror(errno));
        return PPGS_E_INVALID;
    }
    if (!write_all(out.fd, head.b.data(), head.b.size()) || !write_all(out.fd, payload, payload_bytes) ||
        !write_all(out.fd, tail.b.data(), tail.b.size())) {
        set_error("%s: write failed: %s", path, strerror(errno));
        return PPGS_E_INVALID;
    }
    return PPGS_OK;
}

// ---------------------------------------------------------------------------
// .pt reader: one contiguous 1..3-D fp16 / fp32 CPU tensor saved by torch.save (zip container,
// stored entries) — the feature caches `data/cache/<dataset>/<stem>-mel.pt` that
// ppgs/data/dataset.py:98-101 loads with torch.load, and this library's own outputs.
// ---------------------------------------------------------------------------
struct PtInfo {
    int64_t dims[3] = {1, 1, 1};
    int ndim = 0, elem = 0;
    int64_t data_offset = 0, data_bytes = 0;
};


// Walks the pickle of `torch.save(tensor)`: the storage class names the element type, the ints
// after BINPERSID are (storage offset, size tuple, stride tuple).  Only the opcodes torch's
// pickler emits for a plain tensor are understood; anything else is "unsupported" and the
// caller falls back to torch.load.
static int parse_tensor_pickle(const unsigned char* b, size_t n, PtInfo* info, const char* path) {
    std::vector<int64_t> ints;
    std::vector<std::vector<int64_t>> tuples;
    bool after_persid = false, first_object = true;
    int64_t storage_offset = -1;
    size_t i = 0;
    auto need = [&](size_t k) { return i + k <= n; };
    while (i < n) {
        const unsigned char op = b[i++];
        if (first_object && op != 0x80 && op != 0x95) {
            // the pickled object must BE a tensor (not a dict / list that contains one)
            static const char kRebuild[] = "torch._utils\n_rebuild_tensor_v2\n";
            if (op != 'c' || !need(sizeof(kRebuild) - 1) || memcmp(b + i, kRebuild, sizeof(kRebuild) - 1) != 0) {
                set_error("%s: the file does not hold a bare tensor", path);
                return PPGS_E_UNSUPPORTED;
            }
            first_object = false;
        }
        switch (op) {
            case 0x80: if (!need(1)) goto bad; i += 1; break;                         // PROTO
            case 0x95: if (!need(8)) goto bad; i += 8; break;                         // FRAME
            case 'c': {                                                               // GLOBAL module\nname\n
                const size_t start = i;
                int newlines = 0;
                while (i < n && newlines < 2) newlines += b[i++] == '\n';
                if (newlines < 2) goto bad;
                const std::string g(reinterpret_cast<const char*>(b + start), i - start);
                if (g.find("HalfStorage") != std::string::npos) info->elem = 2;
                else if (g.find("FloatStorage") != std::string::npos) info->elem = 4;
                else if (g.find("Storage") != std::string::npos) {
                    set_error("%s: tensor storage %s is not fp16 / fp32", path, g.c_str());
                    return PPGS_E_UNSUPPORTED;
                }
                break;
            }
            case 'X': case 'T': if (!need(4)) goto bad; { const uint32_t len = rd32(b + i); i += 4; if (!need(len)) goto bad; i += len; } break;
            case 0x8c: case 'U': if (!need(1)) goto bad; { const size_t len = b[i]; i += 1; if (!need(len)) goto bad; i += len; } break;
            case 'K': if (!need(1)) goto bad; if (after_persid) ints.push_back(b[i]); i += 1; break;
            case 'M': if (!need(2)) goto bad; if (after_persid) ints.push_back(rd16(b + i)); i += 2; break;
            case 'J': if (!need(4)) goto bad; if (after_persid) ints.push_back((int32_t)rd32(b + i)); i += 4; break;
            case 0x8a: {                                                              // LONG1
                if (!need(1)) goto bad;
                const size_t len = b[i]; i += 1;
                if (!need(len) || len > 8) goto bad;
                int64_t v = 0;
                for (size_t k = 0; k < len; ++k) v |= (int64_t)b[i + k] << (8 * k);
                if (after_persid) ints.push_back(v);
                i += len;
                break;
            }
            case 'q': case 'h': if (!need(1)) goto bad; i += 1; break;               // BINPUT / BINGET
            case 'r': case 'j': if (!need(4)) goto bad; i += 4; break;               // LONG_BINPUT / LONG_BINGET
            case 'Q': after_persid = true; ints.clear(); break;                       // BINPERSID
            case 0x85: case 0x86: case 0x87: {                                        // TUPLE1..3
                const size_t k = op - 0x84;
                if (after_persid && tuples.size() < 2) {
                    if (ints.size() < k) goto bad;
                    if (tuples.empty()) {
                        if (ints.size() != k + 1) goto bad;   // storage offset, then the sizes
                        storage_offset = ints[0];
                    }
                    tuples.emplace_back(ints.end() - k, ints.end());
                    ints.clear();
                }
                break;
            }
            case ')': if (after_persid && tuples.size() < 2) { tuples.emplace_back(); ints.clear(); } break;   // 0-d
            case '(': case 't': case 'R': case '}': case ']': case 'N': case 0x88: case 0x89: case 'b':
            case 'u': case 's': case 0x94: case 'a': case 'e':
                break;
            case '.': i = n; break;
            default:
                set_error("%s: pickle opcode 0x%02x is not part of a plain tensor file", path, op);
                return PPGS_E_UNSUPPORTED;
        }
        if (tuples.size() == 2 && info->ndim == 0 && !tuples[0].empty()) {
            const std::vector<int64_t>&size = tuples[0], &stride = tuples[1];
            if (size.size() != stride.size() || size.size() > 3 || storage_offset != 0) {
                set_error("%s: not a whole contiguous tensor (storage offset %lld)", path, (long long)storage_offset);
                return PPGS_E_UNSUPPORTED;
            }
            int64_t expect = 1;
            for (int d = (int)size.size() - 1; d >= 0; --d) {
                if (size[d] < 0 || (size[d] > 1 && stride[d] != expect)) {
                    set_error("%s: tensor is not contiguous", path);
                    return PPGS_E_UNSUPPORTED;
                }
                expect *= size[d];
            }
            info->ndim = (int)size.size();
            for (size_t d = 0; d < size.size(); ++d) info->dims[3 - size.size() + d] = size[d];
        }
    }
    if (info->ndim == 0 || info->elem == 0) {
        set_error("%s: no fp16 / fp32 tensor found in data.pkl", path);
        return PPGS_E_UNSUPPORTED;
    }
    return PPGS_OK;
bad:
    set_error("%s: truncated or unexpected pickle stream", path);
    return PPGS_E_UNSUPPORTED;
}

static int pt_probe(int fd, const char* path, PtInfo* info) {
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 22) {
        set_error("%s: not a torch.save archive", path);
        return PPGS_E_INVALID;
    }
    const size_t tail = (size_t)std::min<int64_t>(st.st_size, 65536 + 22);
    std::vector<unsigned char> buf(tail);
    if (pread(fd, buf.data(), tail, st.st_size - (int64_t)tail) != (ssize_t)tail) {
        set_error("%s: read failed: %s", path, strerror(errno));
        return PPGS_E_INVALID;
    }
    int64_t eocd = -1;
    for (int64_t i = (int64_t)tail - 22; i >= 0; --i)
        if (rd32(&buf[i]) == 0x06054b50) { eocd = i; break; }
    if (eocd < 0) {
        set_error("%s: not a zip archive (legacy torch.save format is not supported)", path);
        return PPGS_E_UNSUPPORTED;
    }
    const uint32_t entries = rd16(&buf[eocd + 10]), cd_size = rd32(&buf[eocd + 12]), cd_off = rd32(&buf[eocd + 16]);
    if (cd_off == 0xFFFFFFFFu || entries == 0xFFFF) {
        set_error("%s: zip64 archives are not supported by the native reader", path);
        return PPGS_E_UNSUPPORTED;
    }
    std::vector<unsigned char> cd(cd_size);
    if (pread(fd, cd.data(), cd_size, cd_off) != (ssize_t)cd_size) {
        set_error("%s: central directory read failed", path);
        return PPGS_E_INVALID;
    }
    int64_t pkl_local = -1, pkl_size = 0, data_local = -1, data_size = 0;
    size_t at = 0;
    for (uint32_t k = 0; k < entries && at + 46 <= cd.size(); ++k) {
        if (rd32(&cd[at]) != 0x02014b50) break;
        const uint16_t method = rd16(&cd[at + 10]), name_len = rd16(&cd[at + 28]), extra_len = rd16(&cd[at + 30]),
                       comment_len = rd16(&cd[at + 32]);
        const uint32_t comp = rd32(&cd[at + 20]), local = rd32(&cd[at + 42]);
        const std::string name(reinterpret_cast<const char*>(&cd[at + 46]), name_len);
        auto ends_with = [&](const char* suffix) {
            const size_t m = strlen(suffix);
            return name.size() >= m && name.compare(name.size() - m, m, suffix) == 0;
        };
        if (ends_with("/data.pkl") || ends_with("/data/0")) {
            if (method != 0) {
                set_error("%s: entry %s is compressed", path, name.c_str());
                return PPGS_E_UNSUPPORTED;
            }
            if (ends_with("/data.pkl")) { pkl_local = local; pkl_size = comp; }
            else { data_local = local; data_size = comp; }
        } else if (name.find("/data/") != std::string::npos) {
            set_error("%s: more than one tensor storage in the archive", path);
            return PPGS_E_UNSUPPORTED;
        }
        at += 46 + name_len + extra_len + comment_len;
    }
    if (pkl_local < 0 || data_local < 0) {
        set_error("%s: data.pkl / data/0 not found", path);
        return PPGS_E_UNSUPPORTED;
    }
    auto payload_offset = [&](int64_t local, int64_t* out) -> bool {
        unsigned char h[30];
        if (pread(fd, h, 30, local) != 30 || rd32(h) != 0x04034b50) return false;
        *out = local + 30 + rd16(h + 26) + rd16(h + 28);
        return true;
    };
    int64_t pkl_at = 0;
    if (!payload_offset(pkl_local, &pkl_at) || !payload_offset(data_local, &info->data_offset)) {
        set_error("%s: bad local file header", path);
        return PPGS_E_INVALID;
    }
    std::vector<unsigned char> pkl((size_t)pkl_size);
    if (pread(fd, pkl.data(), pkl.size(), pkl_at) != (ssize_t)pkl.size()) {
        set_error("%s: data.pkl read failed", path);
        return PPGS_E_INVALID;
    }
    PPGS_CHECK(parse_tensor_pickle(pkl.data(), pkl.size(), info, path));
    info->data_bytes = info->dims[0] * info->dims[1] * info->dims[2] * info->elem;
    if (info->data_bytes != data_size) {
        set_error("%s: storage holds %lld bytes, the tensor needs %lld", path, (long long)data_size,
                  (long long)info->data_bytes);
        return PPGS_E_UNSUPPORTED;
    }
    return PPGS_OK;
}

// ---------------------------------------------------------------------------
// Device kernels
// ---------------------------------------------------------------------------
// int16 PCM -> fp32 / 32768: 8 samples per thread (16-byte load, two 16-byte stores)
__global__ void __launch_bounds__(256)
pcm16_to_f32_kernel(const int16_t* __restrict__ pcm, float* __restrict__ out, int64_t count) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i + 8 <= count) {
        const int4 raw = *reinterpret_cast<const int4*>(pcm + i);
        const int w[4] = {raw.x, raw.y, raw.z, raw.w};
        float v[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[2 * k] = (float)(int16_t)(w[k] & 0xFFFF) * (1.0f / 32768.0f);
            v[2 * k + 1] = (float)(int16_t)(w[k] >> 16) * (1.0f / 32768.0f);
        }
        *reinterpret_cast<float4*>(out + i) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(out + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        for (int64_t k = i; k < count; ++k) out[k] = (float)pcm[k] * (1.0f / 32768.0f);
    }
}

// out[b][j * up + p] = sum_k x_pad[b][j * down + k] * taps[k][p], x_pad = x shifted by `width`
// zeros (torchaudio _apply_sinc_resample_kernel: pad (width, width + down), conv1d stride
// down).  Block = 32 phases-groups x 8 j; the input segment of the 8 j's is staged in smem;
// `taps` is stored [k][up] so that the phases of one k are contiguous.
constexpr int kResampleJ = 8;
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ x, int64_t in_len, int64_t in_stride, const float* __restrict__ taps,
                int up, int down, int width, int ntaps, float* __restrict__ out, int64_t out_len,
                int64_t out_stride) {
    extern __shared__ float seg[];   // (kResampleJ - 1) * down + ntaps samples
    const int b = blockIdx.y;
    const int64_t j0 = (int64_t)blockIdx.x * kResampleJ;
    const int seg_len = (kResampleJ - 1) * down + ntaps;
    const float* xb = x + (int64_t)b * in_stride;
    const int64_t first = j0 * down - width;
    for (int i = threadIdx.x; i < seg_len; i += blockDim.x) {
        const int64_t at = first + i;
        seg[i] = (at >= 0 && at < in_len) ? xb[at] : 0.f;
    }
    __syncthreads();
    const int jj = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* s = seg + jj * down;
    for (int p = lane; p < up; p += 32) {
        const int64_t o = (j0 + jj) * up + p;
        if (o >= out_len) continue;
        float acc = 0.f;
        for (int k = 0; k < ntaps; ++k) acc = fmaf(s[k], __ldg(taps + (int64_t)k * up + p), acc);
        out[(int64_t)b * out_stride + o] = acc;
    }
}

static int64_t gcd64(int64_t a, int64_t b) {
    while (b) {
        const int64_t t = a % b;
        a = b;
        b = t;
    }
    return a;
}

// torchaudio functional._get_sinc_resample_kernel (sinc_interp_hann, lowpass_filter_width 6,
// rolloff 0.99), evaluated in double and rounded to fp32 like the reference; layout [k][up].
static void resample_taps(int up, int down, std::vector<float>& taps, int* width_out) {
    const int lowpass = 6;
    const double rolloff = 0.99;
    const double base = (double)(up < down ? up : down) * rolloff;
    const int width = (int)ceil((double)lowpass * down / base);
    const int ntaps = 2 * width + down;
    taps.assign((size_t)ntaps * up, 0.f);
    const double scale = base / down;
    for (int p = 0; p < up; ++p)
        for (int k = 0; k < ntaps; ++k) {
            // the reference forms the phase offset -p / up in fp32 (an int64 arange divided by an
            // int promotes to the default dtype) before adding the fp64 tap grid
            const float phase = (float)(-p) / (float)up;
            double t = ((double)phase + (double)(k - width) / down) * base;
            t = t < -lowpass ? -lowpass : (t > lowpass ? lowpass : t);
            const double c = cos(t * M_PI / lowpass / 2.0);
            const double window = c * c;
            const double a = t * M_PI;
            const double sinc = a == 0.0 ? 1.0 : sin(a) / a;
            taps[(size_t)k * up + p] = (float)(sinc * (window * scale));
        }
    *width_out = width;
}

struct ResampleCache {
    std::mutex lock;
    std::map<std::pair<int, std::pair<int, int>>, std::pair<float*, int>> dev;   // (device, up, down)
};
static ResampleCache resample_cache;

// ---------------------------------------------------------------------------
// File pipeline
// ---------------------------------------------------------------------------
struct Pipeline {
    ppgs_engine* e;
    cudaStream_t stream;
    int n_batches, legacy_mode;
    const int32_t* batch_sizes;
    const char* const* audio_files;
    const char* const* output_files;
    const int64_t* file_samples;
    // flat index of a batch's first file, its longest file (samples), row stride of its buffers
    std::vector<int64_t> batch_first, batch_max, batch_stride;

    int n_in, n_out;
    std::vector<int16_t*> in_host;    // pinned [batch][max_samples]
    std::vector<float*> out_host;     // pinned [batch][O][frames]
    size_t in_bytes = 0, out_bytes = 0;

    std::mutex lock;
    std::condition_variable cv;
    std::vector<int> in_pending;      // per batch: files still to read (-1 = slot not granted yet)
    std::vector<int> out_pending;     // per batch: files still to write
    std::atomic<int64_t> next_file{0};
    std::deque<std::pair<int, int>> write_queue;   // (batch, row)
    std::vector<cudaEvent_t> d2h_done;             // per out slot
    bool writers_finish = false;
    int error = PPGS_OK;
    std::string error_text;
    std::atomic<int64_t> frames_done{0};

    void fail(int code) {
        std::lock_guard<std::mutex> g(lock);
        if (error == PPGS_OK) {
            error = code;
            error_text = ppgs_last_error();
        }
        cv.notify_all();
    }
    bool failed() {
        std::lock_guard<std::mutex> g(lock);
        return error != PPGS_OK;
    }
};

static void reader_main(Pipeline* p, int64_t total_files) {
    cudaSetDevice(p->e->device);
    std::vector<int16_t> scratch;
    int batch = 0;
    for (;;) {
        const int64_t f = p->next_file.fetch_add(1);
        if (f >= total_files) return;
        while (batch + 1 < p->n_batches && p->batch_first[batch + 1] <= f) ++batch;
        const int row = (int)(f - p->batch_first[batch]);
        {   // the batch's input slot is granted by the main thread once its previous user is on the device
            std::unique_lock<std::mutex> g(p->lock);
            p->cv.wait(g, [&] { return p->in_pending[batch] >= 0 || p->error != PPGS_OK; });
            if (p->error != PPGS_OK) return;
        }
        const int64_t max_samples = p->batch_stride[batch];
        int16_t* dst = p->in_host[batch % p->n_in] + (int64_t)row * max_samples;
        const char* path = p->audio_files[f];
        int rc = PPGS_OK;
        const size_t path_len = strlen(path);
        if (path_len > 5 && strcasecmp(path + path_len - 5, ".flac") == 0) {
            rc = flac_read_pcm16(path, dst, max_samples, p->file_samples[f]);
        } else {
            Fd in(path, O_RDONLY);
            WavHeader h;
            if (in.fd < 0) {
                set_error("%s: cannot open: %s", path, strerror(errno));
                rc = PPGS_E_INVALID;
            } else if ((rc = parse_wav(in.fd, path, &h)) == PPGS_OK) {
                if (h.is_float || h.bits != 16 || h.sample_rate != 16000 || h.frames != p->file_samples[f]) {
                    set_error("%s: the native file pipeline takes 16-bit PCM at 16 kHz with the announced "
                              "length (got %d bits%s, %d Hz, %lld frames, expected %lld)",
                              path, h.bits, h.is_float ? " float" : "", h.sample_rate, (long long)h.frames,
                              (long long)p->file_samples[f]);
                    rc = PPGS_E_UNSUPPORTED;
                } else {
                    rc = read_pcm16_channel0(in.fd, path, h, h.frames, dst, scratch);
                    if (rc == PPGS_OK && h.frames < max_samples)
                        memset(dst + h.frames, 0, (size_t)(max_samples - h.frames) * 2);
                }
            }
        }
        if (rc != PPGS_OK) {
            p->fail(rc);
            return;
        }
        std::lock_guard<std::mutex> g(p->lock);
        if (--p->in_pending[batch] == 0) p->cv.notify_all();
    }
}

static void writer_main(Pipeline* p) {
    cudaSetDevice(p->e->device);
    std::vector<float> contiguous;
    ByteSink head;
    const int O = p->e->cfg.output_channels;
    for (;;) {
        std::pair<int, int> item;
        {
            std::unique_lock<std::mutex> g(p->lock);
            p->cv.wait(g, [&] { return !p->write_queue.empty() || p->writers_finish || p->error != PPGS_OK; });
            if (p->error != PPGS_OK) return;
            if (p->write_queue.empty()) return;   // writers_finish
            item = p->write_queue.front();
            p->write_queue.pop_front();
        }
        const int batch = item.first, row = item.second;
        if (cudaEventSynchronize(p->d2h_done[batch % p->n_out]) != cudaSuccess) {
            set_error("file pipeline: waiting for the D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            p->fail(PPGS_E_CUDA);
            return;
        }
        const int64_t f = p->batch_first[batch] + row;
        const int64_t frames_max = p->batch_max[batch] / kHopSamples;
        const int64_t frames = p->file_samples[f] / kHopSamples;
        const float* src = p->out_host[batch % p->n_out] + (int64_t)row * O * frames_max;
        const int rc = pt_write(p->output_files[f], src, O, frames, frames_max, contiguous, head);
        if (rc != PPGS_OK) {
            p->fail(rc);
            return;
        }
        p->frames_done.fetch_add(frames);
        std::lock_guard<std::mutex> g(p->lock);
        if (--p->out_pending[batch] == 0) p->cv.notify_all();
    }
}

}  // namespace ppgs

using namespace ppgs;

struct IoDeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit IoDeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) {
            set_error("cannot select CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
            ok = false;
        }
    }
    ~IoDeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

extern "C" {

int ppgs_wav_info(const char* path, int64_t* frames, int* sample_rate, int* channels, int* bits,
                  int* is_float) {
    if (!path) {
        set_error("wav_info: path is NULL");
        return PPGS_E_INVALID;
    }
    Fd in(path, O_RDONLY);
    if (in.fd < 0) {
        set_error("%s: cannot open: %s", path, strerror(errno));
        return PPGS_E_INVALID;
    }
    WavHeader h;
    PPGS_CHECK(parse_wav(in.fd, path, &h));
    if (frames) *frames = h.frames;
    if (sample_rate) *sample_rate = h.sample_rate;
    if (channels) *channels = h.channels;
    if (bits) *bits = h.bits;
    if (is_float) *is_float = h.is_float;
    return PPGS_OK;
}

int ppgs_wav_info_many(const char* const* paths, int64_t count, int threads, int64_t* frames,
                       int32_t* sample_rate, int32_t* channels, int32_t* bits, int32_t* is_float,
                       int32_t* status) {
    if (!paths || count < 0 || !frames || !sample_rate || !channels || !bits || !is_float || !status) {
        set_error("wav_info_many: bad argument");
        return PPGS_E_INVALID;
    }
    threads = threads < 1 ? 1 : (threads > 64 ? 64 : threads);
    if (count < 64) threads = 1;
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= count) return;
            int rate = 0, ch = 0, b = 0, f = 0;
            int64_t n = 0;
            status[i] = ppgs_wav_info(paths[i], &n, &rate, &ch, &b, &f);
            frames[i] = n;
            sample_rate[i] = rate;
            channels[i] = ch;
            bits[i] = b;
            is_float[i] = f;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return PPGS_OK;
}

int ppgs_wav_read_f32(const char* path, float* dst, int64_t capacity, int64_t* frames, int* sample_rate) {
    if (!path || !dst) {
        set_error("wav_read_f32: bad argument");
        return PPGS_E_INVALID;
    }
    Fd in(path, O_RDONLY);
    if (in.fd < 0) {
        set_error("%s: cannot open: %s", path, strerror(errno));
        return PPGS_E_INVALID;
    }
    WavHeader h;
    PPGS_CHECK(parse_wav(in.fd, path, &h));
    if (h.frames > capacity) {
        set_error("%s: %lld frames do not fit the buffer of %lld", path, (long long)h.frames,
                  (long long)capacity);
        return PPGS_E_INVALID;
    }
    std::vector<unsigned char> raw((size_t)h.data_bytes);
    if (h.data_bytes && !pread_all(in.fd, raw.data(), raw.size(), h.data_offset)) {
        set_error("%s: short read", path);
        return PPGS_E_INVALID;
    }
    PPGS_CHECK(decode_channel0(h, raw.data(), h.frames, dst));
    if (frames) *frames = h.frames;
    if (sample_rate) *sample_rate = h.sample_rate;
    return PPGS_OK;
}

int ppgs_pt_write_f32(const char* path, const float* data, int64_t rows, int64_t cols, int64_t row_stride) {
    if (!path || (!data && rows * cols > 0)) {
        set_error("pt_write_f32: bad argument");
        return PPGS_E_INVALID;
    }
    std::vector<float> contiguous;
    ByteSink head;
    return pt_write(path, data, rows, cols, row_stride, contiguous, head);
}

int ppgs_pt_write_f16(const char* path, const void* data, int64_t rows, int64_t cols, int64_t row_stride) {
    if (!path || (!data && rows * cols > 0)) {
        set_error("pt_write_f16: bad argument");
        return PPGS_E_INVALID;
    }
    std::vector<float> contiguous;
    ByteSink head;
    return pt_write(path, data, rows, cols, row_stride, contiguous, head, 2);
}

int ppgs_pt_info(const char* path, int* ndim, int64_t* dims3, int* elem_bytes) {
    if (!path || !ndim || !dims3 || !elem_bytes) {
        set_error("pt_info: NULL argument");
        return PPGS_E_INVALID;
    }
    Fd in(path, O_RDONLY);
    if (in.fd < 0) {
        set_error("%s: cannot open: %s", path, strerror(errno));
        return PPGS_E_INVALID;
    }
    PtInfo info;
    PPGS_CHECK(pt_probe(in.fd, path, &info));
    *ndim = info.ndim;
    for (int d = 0; d < 3; ++d) dims3[d] = info.dims[d];
    *elem_bytes = info.elem;
    return PPGS_OK;
}

int ppgs_pt_read(const char* path, void* dst_host, int64_t rows, int64_t cols, int elem_bytes,
                 int64_t dst_row_stride) {
    if (!path || !dst_host || rows < 0 || cols < 0 || dst_row_stride < cols) {
        set_error("pt_read: bad argument");
        return PPGS_E_INVALID;
    }
    Fd in(path, O_RDONLY);
    if (in.fd < 0) {
        set_error("%s: cannot open: %s", path, strerror(errno));
        return PPGS_E_INVALID;
    }
    PtInfo info;
    PPGS_CHECK(pt_probe(in.fd, path, &info));
    if (info.elem != elem_bytes || info.dims[0] * info.dims[1] != rows || info.dims[2] != cols) {
        set_error("%s: holds (%lld, %lld) x %d bytes, the caller expects (%lld, %lld) x %d", path,
                  (long long)(info.dims[0] * info.dims[1]), (long long)info.dims[2], info.elem, (long long)rows,
                  (long long)cols, elem_bytes);
        return PPGS_E_INVALID;
    }
    char* dst = static_cast<char*>(dst_host);
    const size_t row_bytes = (size_t)cols * elem_bytes;
    if (dst_row_stride == cols) {
        if (pread(in.fd, dst, row_bytes * rows, info.data_offset) != (ssize_t)(row_bytes * rows)) {
            set_error("%s: payload read failed: %s", path, strerror(errno));
            return PPGS_E_INVALID;
        }
        return PPGS_OK;
    }
    for (int64_t r = 0; r < rows; ++r)   // rows land in a padded batch tensor
        if (pread(in.fd, dst + r * dst_row_stride * elem_bytes, row_bytes, info.data_offset + r * (int64_t)row_bytes) !=
            (ssize_t)row_bytes) {
            set_error("%s: payload read failed: %s", path, strerror(errno));
            return PPGS_E_INVALID;
        }
    return PPGS_OK;
}

int64_t ppgs_resample_length(int64_t samples, int orig_rate, int target_rate) {
    if (samples < 0 || orig_rate <= 0 || target_rate <= 0) return -1;
    const int64_t g = gcd64(orig_rate, target_rate);
    const int64_t up = target_rate / g, down = orig_rate / g;
    return (up * samples + down - 1) / down;   // ceil(new * length / orig)
}

int ppgs_resample_taps(int orig_rate, int target_rate, float* taps, int64_t capacity, int* ntaps, int* phases,
                       int* width) {
    if (orig_rate <= 0 || target_rate <= 0) {
        set_error("resample: rates must be positive integers");
        return PPGS_E_INVALID;
    }
    const int64_t g = gcd64(orig_rate, target_rate);
    const int up = (int)(target_rate / g), down = (int)(orig_rate / g);
    std::vector<float> table;
    int w = 0;
    resample_taps(up, down, table, &w);
    if (ntaps) *ntaps = 2 * w + down;
    if (phases) *phases = up;
    if (width) *width = w;
    if (taps) {
        if ((int64_t)table.size() > capacity) {
            set_error("resample_taps: table of %zu floats does not fit %lld", table.size(), (long long)capacity);
            return PPGS_E_INVALID;
        }
        memcpy(taps, table.data(), table.size() * 4);
    }
    return PPGS_OK;
}

int ppgs_resample(ppgs_engine* e, const float* audio_dev, int batch, int64_t samples, int64_t audio_stride,
                  int orig_rate, int target_rate, float* out_dev, int64_t out_stride, void* stream_) {
    if (!e) {
        set_error("engine is NULL");
        return PPGS_E_INVALID;
    }
    IoDeviceGuard guard(e->device);
    if (!guard.ok) return PPGS_E_CUDA;
    if (!audio_dev || !out_dev || batch <= 0 || samples <= 0 || orig_rate <= 0 || target_rate <= 0 ||
        audio_stride < samples) {
        set_error("resample: bad argument");
        return PPGS_E_INVALID;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t out_len = ppgs_resample_length(samples, orig_rate, target_rate);
    if (out_stride < out_len) {
        set_error("resample: output stride %lld < %lld output samples", (long long)out_stride, (long long)out_len);
        return PPGS_E_INVALID;
    }
    const int64_t g = gcd64(orig_rate, target_rate);
    const int up = (int)(target_rate / g), down = (int)(orig_rate / g);
    float* taps_dev = nullptr;
    int width = 0;
    {
        std::lock_guard<std::mutex> lock(resample_cache.lock);
        auto key = std::make_pair(e->device, std::make_pair(up, down));
        auto it = resample_cache.dev.find(key);
        if (it == resample_cache.dev.end()) {
            std::vector<float> table;
            resample_taps(up, down, table, &width);
            PPGS_CUDA(cudaMalloc(&taps_dev, table.size() * 4));
            PPGS_CUDA(cudaMemcpy(taps_dev, table.data(), table.size() * 4, cudaMemcpyHostToDevice));
            resample_cache.dev[key] = std::make_pair(taps_dev, width);
        } else {
            taps_dev = it->second.first;
            width = it->second.second;
        }
    }
    const int ntaps = 2 * width + down;
    const size_t smem = (size_t)((kResampleJ - 1) * down + ntaps) * 4;
    if (smem > 200 * 1024) {
        set_error("resample: rate ratio %d:%d needs %zu bytes of shared memory; reduce the ratio", up, down, smem);
        return PPGS_E_UNSUPPORTED;
    }
    PPGS_CUDA(cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t groups = (out_len + up - 1) / up;
    dim3 grid((unsigned)((groups + kResampleJ - 1) / kResampleJ), (unsigned)batch);
    {
        LaunchScope scope(e, "resample_sinc", stream);
        resample_kernel<<<grid, 256, smem, stream>>>(audio_dev, samples, audio_stride, taps_dev, up, down, width,
                                                     ntaps, out_dev, out_len, out_stride);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

int ppgs_pcm16_to_f32(ppgs_engine* e, const void* pcm_dev, int64_t count, float* out_dev, void* stream_) {
    if (!e) {
        set_error("engine is NULL");
        return PPGS_E_INVALID;
    }
    IoDeviceGuard guard(e->device);
    if (!guard.ok) return PPGS_E_CUDA;
    if (!pcm_dev || !out_dev || count <= 0 || ((uintptr_t)pcm_dev & 15) || ((uintptr_t)out_dev & 15)) {
        set_error("pcm16_to_f32: bad argument (buffers must be 16-byte aligned)");
        return PPGS_E_INVALID;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t threads = (count + 7) / 8;
    {
        LaunchScope scope(e, "pcm16_to_f32", stream);
        pcm16_to_f32_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
            static_cast<const int16_t*>(pcm_dev), out_dev, count);
    }
    PPGS_CUDA(cudaGetLastError());
    return PPGS_OK;
}

int ppgs_files_to_files(ppgs_engine* e, int n_batches, const int32_t* batch_sizes,
                        const char* const* audio_files, const char* const* output_files,
                        const int64_t* file_samples, int reader_threads, int writer_threads, int legacy_mode,
                        void* stream_, int64_t* frames_done) {
    if (!e) {
        set_error("engine is NULL");
        return PPGS_E_INVALID;
    }
    IoDeviceGuard guard(e->device);
    if (!guard.ok) return PPGS_E_CUDA;
    if (!e->finalized) {
        set_error("engine has no weights: call ppgs_engine_finalize first");
        return PPGS_E_STATE;
    }
    if (frames_done) *frames_done = 0;
    if (n_batches == 0) return PPGS_OK;
    if (n_batches < 0 || !batch_sizes || !audio_files || !output_files || !file_samples) {
        set_error("files_to_files: bad argument");
        return PPGS_E_INVALID;
    }
    if (e->cfg.input_channels != kMelChannels) {
        set_error("files_to_files: the native pipeline runs the mel front-end; model expects %d channels",
                  e->cfg.input_channels);
        return PPGS_E_INVALID;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    reader_threads = reader_threads < 1 ? 1 : (reader_threads > 64 ? 64 : reader_threads);
    writer_threads = writer_threads < 1 ? 1 : (writer_threads > 64 ? 64 : writer_threads);
    const int O = e->cfg.output_channels;

    Pipeline p;
    p.e = e;
    p.stream = stream;
    p.n_batches = n_batches;
    p.legacy_mode = legacy_mode;
    p.batch_sizes = batch_sizes;
    p.audio_files = audio_files;
    p.output_files = output_files;
    p.file_samples = file_samples;
    p.batch_first.resize(n_batches);
    p.batch_max.resize(n_batches);
    p.batch_stride.resize(n_batches);
    int64_t total_files = 0;
    size_t audio_dev_bytes = 0, mel_bytes = 0;
    for (int b = 0; b < n_batches; ++b) {
        if (batch_sizes[b] <= 0) {
            set_error("files_to_files: batch %d is empty", b);
            return PPGS_E_INVALID;
        }
        p.batch_first[b] = total_files;
        int64_t longest = 0;
        for (int r = 0; r < batch_sizes[b]; ++r) {
            const int64_t n = file_samples[total_files + r];
            longest = n > longest ? n : longest;
        }
        if (longest < 433) {
            set_error("batch %d: longest file has %lld samples; the mel front-end needs at least 433 "
                      "(reflection padding of 432, as torch does)", b, (long long)longest);
            return PPGS_E_INVALID;
        }
        p.batch_max[b] = longest;
        // 16-byte rows for the vector loads of the decode kernel; the tail stays zero
        const int64_t pitch = (longest + 7) & ~int64_t(7);
        p.batch_stride[b] = pitch;
        const size_t count = (size_t)batch_sizes[b] * pitch;
        p.in_bytes = count * 2 > p.in_bytes ? count * 2 : p.in_bytes;
        const size_t out_b = (size_t)batch_sizes[b] * O * (longest / kHopSamples) * 4;
        p.out_bytes = out_b > p.out_bytes ? out_b : p.out_bytes;
        audio_dev_bytes = count * 4 > audio_dev_bytes ? count * 4 : audio_dev_bytes;
        const size_t mel_b = (size_t)batch_sizes[b] * kMelChannels * (longest / kHopSamples) * 2;
        mel_bytes = mel_b > mel_bytes ? mel_b : mel_bytes;
        total_files += batch_sizes[b];
    }
    auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
    p.n_in = reader_threads < 3 ? 3 : 4;
    p.n_out = 3;
    if (p.n_in > n_batches) p.n_in = n_batches;
    if (p.n_out > n_batches) p.n_out = n_batches;

    // resources (released on every path by the guard below)
    struct Resources {
        std::vector<void*> pinned, device;
        std::vector<cudaEvent_t> events;
        std::vector<cudaStream_t> streams;
        ~Resources() {
            for (cudaStream_t s : streams) cudaStreamDestroy(s);
            for (cudaEvent_t ev : events) cudaEventDestroy(ev);
            for (void* ptr : pinned) cudaFreeHost(ptr);
            for (void* ptr : device) cudaFree(ptr);
        }
    } res;
    auto pinned = [&](size_t bytes, void** out) -> int {
        PPGS_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
        res.pinned.push_back(*out);
        return PPGS_OK;
    };
    auto event = [&](cudaEvent_t* out, bool blocking = false) -> int {
        // blocking sync for the events the WRITER threads wait on: they sleep in
        // cudaEventSynchronize instead of spinning — eight spinning writers per GPU starve the
        // readers and the launch thread as soon as ranks share a host (measured: 2 ranks on 24
        // cores ran at half speed).  The pipeline thread keeps spinning on its own events: its
        // wake-up latency is GPU idle time.
        PPGS_CUDA(cudaEventCreateWithFlags(out, cudaEventDisableTiming | (blocking ? cudaEventBlockingSync : 0)));
        res.events.push_back(*out);
        return PPGS_OK;
    };
    for (int i = 0; i < p.n_in; ++i) {
        void* ptr = nullptr;
        PPGS_CHECK(pinned(p.in_bytes, &ptr));
        p.in_host.push_back(static_cast<int16_t*>(ptr));
    }
    for (int i = 0; i < p.n_out; ++i) {
        void* ptr = nullptr;
        PPGS_CHECK(pinned(p.out_bytes, &ptr));
        p.out_host.push_back(static_cast<float*>(ptr));
    }
    // two device slots: [pcm int16 | audio fp32 | mel fp16 | posteriors fp32]
    const size_t dev_slot = align(p.in_bytes) + align(audio_dev_bytes) + align(mel_bytes) + align(p.out_bytes);
    char* dev_slots[2] = {nullptr, nullptr};
    const int n_dev = n_batches > 1 ? 2 : 1;
    for (int i = 0; i < n_dev; ++i) {
        void* ptr = nullptr;
        PPGS_CUDA(cudaMalloc(&ptr, dev_slot));
        res.device.push_back(ptr);
        dev_slots[i] = static_cast<char*>(ptr);
    }
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    PPGS_CUDA(cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking));
    res.streams.push_back(copy_in);
    PPGS_CUDA(cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking));
    res.streams.push_back(copy_out);
    std::vector<cudaEvent_t> h2d_done(p.n_in), compute_done(2), dev_free(2);
    p.d2h_done.resize(p.n_out);
    for (auto& ev : h2d_done) PPGS_CHECK(event(&ev));
    for (auto& ev : compute_done) PPGS_CHECK(event(&ev));
    for (auto& ev : dev_free) PPGS_CHECK(event(&ev));
    for (auto& ev : p.d2h_done) PPGS_CHECK(event(&ev, true));

    p.in_pending.assign(n_batches, -1);
    p.out_pending.assign(n_batches, 0);
    for (int b = 0; b < p.n_in; ++b) p.in_pending[b] = batch_sizes[b];   // first slots are free

    std::vector<std::thread> readers, writers;
    for (int i = 0; i < reader_threads; ++i) readers.emplace_back(reader_main, &p, total_files);
    for (int i = 0; i < writer_threads; ++i) writers.emplace_back(writer_main, &p);

    int rc = PPGS_OK;
    auto cuda_ok = [&](cudaError_t err, const char* what) {
        if (err == cudaSuccess) return true;
        set_error("file pipeline: %s failed: %s", what, cudaGetErrorString(err));
        rc = PPGS_E_CUDA;
        return false;
    };
    std::vector<int64_t> lengths;
    for (int b = 0; b < n_batches && rc == PPGS_OK; ++b) {
        const int B = batch_sizes[b];
        const int64_t max_samples = p.batch_max[b], stride = p.batch_stride[b];
        const int64_t frames_max = max_samples / kHopSamples;
        {   // wait: all files of this batch read; the output slot's previous batch written
            std::unique_lock<std::mutex> g(p.lock);
            p.cv.wait(g, [&] {
                return p.error != PPGS_OK ||
                       (p.in_pending[b] == 0 && (b < p.n_out || p.out_pending[b - p.n_out] == 0));
            });
            if (p.error != PPGS_OK) break;
        }
        char* slot = dev_slots[b % n_dev];
        int16_t* pcm_dev = reinterpret_cast<int16_t*>(slot);
        float* audio_dev = reinterpret_cast<float*>(slot + align(p.in_bytes));
        __half* mel_dev = reinterpret_cast<__half*>(slot + align(p.in_bytes) + align(audio_dev_bytes));
        float* out_dev = reinterpret_cast<float*>(slot + align(p.in_bytes) + align(audio_dev_bytes) + align(mel_bytes));
        const size_t count = (size_t)B * stride;
        // the device slot was last used by batch b - 2: its D2H must have drained
        if (b >= n_dev && !cuda_ok(cudaStreamWaitEvent(copy_in, dev_free[b % n_dev], 0), "cudaStreamWaitEvent")) break;
        if (!cuda_ok(cudaMemcpyAsync(pcm_dev, p.in_host[b % p.n_in], count * 2, cudaMemcpyHostToDevice, copy_in),
                     "H2D copy")) break;
        if (!cuda_ok(cudaEventRecord(h2d_done[b % p.n_in], copy_in), "cudaEventRecord")) break;
        if (!cuda_ok(cudaStreamWaitEvent(stream, h2d_done[b % p.n_in], 0), "cudaStreamWaitEvent")) break;
        lengths.assign(B, 0);
        for (int r = 0; r < B; ++r) lengths[r] = file_samples[p.batch_first[b] + r];
        {
            const int64_t threads = ((int64_t)count + 7) / 8;
            LaunchScope scope(e, "pcm16_to_f32", stream);
            pcm16_to_f32_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(pcm_dev, audio_dev,
                                                                                      (int64_t)count);
        }
        if (!cuda_ok(cudaGetLastError(), "pcm16_to_f32 launch")) break;
        // PPGS_B200_FILES_NULL_GPU=1 (measurement only): skip the mel + Transformer kernels and keep
        // everything else (file reads, H2D, D2H of the unwritten buffer, crop, .pt writes) — the
        // host-side ceiling of this pipeline on a given box
        static const bool null_gpu = [] { const char* v = getenv("PPGS_B200_FILES_NULL_GPU"); return v && atoi(v) != 0; }();
        if (!null_gpu)
            rc = ppgs_detail_from_audio_device(e, audio_dev, B, max_samples, stride, lengths.data(), 1, legacy_mode,
                                               out_dev, mel_dev, stream);
        if (rc != PPGS_OK) break;
        if (!cuda_ok(cudaEventRecord(compute_done[b % 2], stream), "cudaEventRecord")) break;
        if (!cuda_ok(cudaStreamWaitEvent(copy_out, compute_done[b % 2], 0), "cudaStreamWaitEvent")) break;
        if (!cuda_ok(cudaMemcpyAsync(p.out_host[b % p.n_out], out_dev, (size_t)B * O * frames_max * 4,
                                     cudaMemcpyDeviceToHost, copy_out), "D2H copy")) break;
        if (!cuda_ok(cudaEventRecord(p.d2h_done[b % p.n_out], copy_out), "cudaEventRecord")) break;
        if (!cuda_ok(cudaEventRecord(dev_free[b % n_dev], copy_out), "cudaEventRecord")) break;
        // the input slot is reusable once its H2D copy has landed: grant it to batch b + n_in
        if (!cuda_ok(cudaEventSynchronize(h2d_done[b % p.n_in]), "cudaEventSynchronize")) break;
        {
            std::lock_guard<std::mutex> g(p.lock);
            if (b + p.n_in < n_batches) p.in_pending[b + p.n_in] = batch_sizes[b + p.n_in];
            p.out_pending[b] = B;
            for (int r = 0; r < B; ++r) p.write_queue.emplace_back(b, r);
        }
        p.cv.notify_all();
    }
    if (rc != PPGS_OK) p.fail(rc);
    {
        std::lock_guard<std::mutex> g(p.lock);
        p.writers_finish = true;
    }
    p.cv.notify_all();
    for (auto& t : readers) t.join();
    for (auto& t : writers) t.join();
    cudaStreamSynchronize(copy_out);
    cudaStreamSynchronize(stream);
    if (frames_done) *frames_done = p.frames_done.load();
    if (p.error != PPGS_OK) {
        set_error("%s", p.error_text.c_str());
        return p.error;
    }
    return PPGS_OK;
}

}  // extern "C"
