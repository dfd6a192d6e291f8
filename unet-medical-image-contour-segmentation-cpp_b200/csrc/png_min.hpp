// png_min.hpp -- dependency-free PNG writer for the reference's side artefacts.
//
// The reference writes `_normalized.png` and `_mask.png` with cv::imwrite(..., PNG_COMPRESSION 0)
// (/root/reference/src/preprocess.cpp:122, src/process.cpp:236-239) and the BGR overlay with default
// compression (src/mask2polygon.cpp:126).  PNG is lossless, so parity is defined on decoded pixels,
// not on file bytes; this writer emits valid PNGs with stored (uncompressed) deflate blocks.
// 8-bit grayscale (channels = 1) or 8-bit RGB (channels = 3, caller passes RGB order).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define MS_PNG_X86 1
#endif

namespace ms {
namespace png {

// CRC-32 (IEEE), slicing-by-8: the artefact writers checksum ~1.3 MB per slice, so the byte-at-a-time form was a
// measurable part of the file path (tools/dir_throughput.py).
struct CrcTables {
    uint32_t t[8][256];
    CrcTables() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int k = 1; k < 8; ++k) t[k][i] = t[0][t[k - 1][i] & 0xFF] ^ (t[k - 1][i] >> 8);
    }
};
inline uint32_t crc32_table(uint32_t crc, const uint8_t* p, size_t n) {
    static const CrcTables T;   // thread-safe initialisation (C++11 magic static)
    crc = ~crc;
    while (n >= 8) {
        uint32_t lo, hi;
        std::memcpy(&lo, p, 4);
        std::memcpy(&hi, p + 4, 4);
        lo ^= crc;              // little-endian host (x86-64 / aarch64)
        crc = T.t[7][lo & 0xFF] ^ T.t[6][(lo >> 8) & 0xFF] ^ T.t[5][(lo >> 16) & 0xFF] ^ T.t[4][lo >> 24] ^
              T.t[3][hi & 0xFF] ^ T.t[2][(hi >> 8) & 0xFF] ^ T.t[1][(hi >> 16) & 0xFF] ^ T.t[0][hi >> 24];
        p += 8;
        n -= 8;
    }
    for (size_t i = 0; i < n; ++i) crc = T.t[0][(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}
#ifdef MS_PNG_X86
// Carry-less-multiply folding (Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ"): 64 bytes per
// iteration in four 128-bit lanes.  `crc` is the running (already inverted) register; len >= 64 and a multiple of 16.
__attribute__((target("pclmul,sse4.1"))) inline uint32_t crc32_clmul(uint32_t crc, const uint8_t* buf, size_t len) {
    alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};
    alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};
    alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};
    alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
    x2 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
    x3 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
    x4 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = _mm_load_si128((const __m128i*)k1k2);
    buf += 64;
    len -= 64;
    while (len >= 64) {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
        x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
        x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
        y6 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
        y7 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
        y8 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
        x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
        x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64;
        len -= 64;
    }
    x0 = _mm_load_si128((const __m128i*)k3k4);      // fold the four lanes into one
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128((const __m128i*)buf);
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16;
        len -= 16;
    }
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);         // 128 -> 64 bits
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64((const __m128i*)k5k0);
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_load_si128((const __m128i*)poly);       // Barrett reduction to 32 bits
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
inline bool cpu_has_clmul() {
    static const bool v = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    return v;
}
inline bool cpu_has_ssse3() {
    static const bool v = __builtin_cpu_supports("ssse3");
    return v;
}
#endif

// CRC-32 (IEEE) of p[0..n) continuing from `crc` (0 to start): PCLMULQDQ folding where the CPU has it, slicing-by-8
// tables otherwise and for the tail.  MEDSEG_SCALAR_CHECKSUMS (read by the tests) forces the table form.
inline bool& force_scalar_checksums() {
    static bool v = std::getenv("MEDSEG_SCALAR_CHECKSUMS") != nullptr;
    return v;
}
inline uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
#ifdef MS_PNG_X86
    if (n >= 64 && cpu_has_clmul() && !force_scalar_checksums()) {
        const size_t body = n & ~(size_t)15;
        crc = ~crc32_clmul(~crc, p, body);
        p += body;
        n -= body;
    }
#endif
    return n ? crc32_table(crc, p, n) : crc;
}

// Adler-32 as two running sums, so scanlines can be fed one at a time: start from {1, 0}, finish with value().
struct Adler {
    uint32_t a = 1, b = 0;
    uint32_t value() const { return (b << 16) | a; }
};
// scalar form: the modulo deferred to once per 5552 bytes (the largest run that cannot overflow 32 bits)
inline void adler_scalar(Adler& s, const uint8_t* p, size_t n) {
    uint32_t a = s.a, b = s.b;
    while (n > 0) {
        const size_t k = n < 5552 ? n : 5552;
        for (size_t i = 0; i < k; ++i) {
            a += p[i];
            b += a;
        }
        a %= 65521u;
        b %= 65521u;
        p += k;
        n -= k;
    }
    s.a = a;
    s.b = b;
}
#ifdef MS_PNG_X86
// 32 bytes per step.  Over a block of m steps starting from (a0, b0):
//   a = a0 + sum(bytes),   b = b0 + 32 m a0 + 32 * sum_j(bytes before step j) + sum_j sum_i (32 - i) * byte[j][i]
// all three sums fit 32-bit lanes for blocks of <= 5536 bytes; the modulo once per block.
__attribute__((target("ssse3"))) inline void adler_ssse3(Adler& s, const uint8_t* p, size_t n) {
    uint64_t a = s.a, b = s.b;
    const __m128i w_hi = _mm_setr_epi8(32, 31, 30, 29, 28, 27, 26, 25, 24, 23, 22, 21, 20, 19, 18, 17);
    const __m128i w_lo = _mm_setr_epi8(16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1);
    const __m128i zero = _mm_setzero_si128(), ones = _mm_set1_epi16(1);
    while (n >= 32) {
        const size_t k = n < 5536 ? (n & ~(size_t)31) : 5536;      // 5536 = 173 * 32
        const uint64_t m = k / 32;
        __m128i va = zero, vs = zero, vb = zero;
        for (size_t i = 0; i < k; i += 32) {
            const __m128i d0 = _mm_loadu_si128((const __m128i*)(p + i)), d1 = _mm_loadu_si128((const __m128i*)(p + i + 16));
            vs = _mm_add_epi32(vs, va);
            va = _mm_add_epi32(va, _mm_add_epi32(_mm_sad_epu8(d0, zero), _mm_sad_epu8(d1, zero)));
            vb = _mm_add_epi32(vb, _mm_madd_epi16(_mm_maddubs_epi16(d0, w_hi), ones));
            vb = _mm_add_epi32(vb, _mm_madd_epi16(_mm_maddubs_epi16(d1, w_lo), ones));
        }
        alignas(16) uint32_t ta[4], ts[4], tb[4];
        _mm_store_si128((__m128i*)ta, va);
        _mm_store_si128((__m128i*)ts, vs);
        _mm_store_si128((__m128i*)tb, vb);
        const uint64_t sum = (uint64_t)ta[0] + ta[2];              // sad_epu8 leaves its sums in lanes 0 and 2
        const uint64_t before = (uint64_t)ts[0] + ts[2];
        const uint64_t weighted = (uint64_t)tb[0] + tb[1] + tb[2] + tb[3];
        b = (b + 32 * m * a + 32 * before + weighted) % 65521u;
        a = (a + sum) % 65521u;
        p += k;
        n -= k;
    }
    s.a = (uint32_t)a;
    s.b = (uint32_t)b;
    adler_scalar(s, p, n);
}
#endif
inline void adler_update(Adler& s, const uint8_t* p, size_t n) {
#ifdef MS_PNG_X86
    if (n >= 64 && cpu_has_ssse3() && !force_scalar_checksums()) return adler_ssse3(s, p, n);
#endif
    adler_scalar(s, p, n);
}
inline uint32_t adler32(const uint8_t* p, size_t n) {
    Adler s;
    adler_update(s, p, n);
    return s.value();
}

inline void put32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}
inline void store32(uint8_t* p, uint32_t x) {
    p[0] = (uint8_t)(x >> 24); p[1] = (uint8_t)(x >> 16); p[2] = (uint8_t)(x >> 8); p[3] = (uint8_t)x;
}

inline void chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& data) {
    put32(out, (uint32_t)data.size());
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    put32(out, crc32_update(0, out.data() + start, out.size() - start));
}

// Encoder.  `fill_row(y, dst)` writes scanline y (w * channels bytes) -- the pixels never exist as a second image: the
// normalised slice is copied, the mask goes through its LUT, the overlay is expanded grey -> RGB with the contour
// pixels painted, all row by row, straight into the zlib stream of stored blocks (which is pure layout: the PNG row
// filter byte 0, then the row; a 5-byte block header every 65,535 stream bytes).
template <class FillRow>
inline std::vector<uint8_t> encode_rows(int w, int h, int channels, FillRow&& fill_row) {
    const size_t row = (size_t)w * channels, raw_n = (row + 1) * (size_t)h;
    const size_t n_blocks = raw_n == 0 ? 1 : (raw_n + 65534) / 65535;
    const size_t z_n = 2 + n_blocks * 5 + raw_n + 4;
    std::vector<uint8_t> out(8 + 25 + 12 + z_n + 12);
    uint8_t* o = out.data();
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::memcpy(o, sig, 8);
    o += 8;
    store32(o, 13);                                      // IHDR
    std::memcpy(o + 4, "IHDR", 4);
    store32(o + 8, (uint32_t)w);
    store32(o + 12, (uint32_t)h);
    o[16] = 8; o[17] = channels == 3 ? 2 : 0; o[18] = 0; o[19] = 0; o[20] = 0;
    store32(o + 21, crc32_update(0, o + 4, 17));
    o += 25;
    uint8_t* idat = o;                                   // IDAT = length | "IDAT" | 78 01 | blocks | adler | crc
    store32(idat, (uint32_t)z_n);
    std::memcpy(idat + 4, "IDAT", 4);
    uint8_t* z = idat + 8;
    *z++ = 0x78;
    *z++ = 0x01;
    std::vector<uint8_t> line(row + 1);
    line[0] = 0;                                         // filter type 0
    Adler ad;
    size_t pos = 0, room = 0, blk = 0;                   // stream position, bytes left in the current stored block
    for (int y = 0; y < h; ++y) {
        fill_row(y, line.data() + 1);
        adler_update(ad, line.data(), row + 1);
        const uint8_t* src = line.data();
        size_t left = row + 1;
        while (left) {
            if (room == 0) {
                const size_t n = std::min<size_t>(65535, raw_n - pos);
                *z++ = ++blk == n_blocks ? 1 : 0;
                *z++ = (uint8_t)(n & 0xFF);
                *z++ = (uint8_t)(n >> 8);
                *z++ = (uint8_t)(~n & 0xFF);
                *z++ = (uint8_t)((~n >> 8) & 0xFF);
                room = n;
            }
            const size_t k = std::min(left, room);
            std::memcpy(z, src, k);
            z += k; src += k; left -= k; room -= k; pos += k;
        }
    }
    if (raw_n == 0) { *z++ = 1; *z++ = 0; *z++ = 0; *z++ = 0xFF; *z++ = 0xFF; }
    store32(z, ad.value());
    z += 4;
    store32(z, crc32_update(0, idat + 4, 4 + z_n));
    z += 4;
    store32(z, 0);                                       // IEND
    std::memcpy(z + 4, "IEND", 4);
    store32(z + 8, crc32_update(0, z + 4, 4));
    return out;
}

inline std::vector<uint8_t> encode(const uint8_t* pixels, int w, int h, int channels) {
    const size_t row = (size_t)w * channels;
    return encode_rows(w, h, channels, [&](int y, uint8_t* dst) { std::memcpy(dst, pixels + (size_t)y * row, row); });
}

inline bool write_bytes(const std::string& path, const std::vector<uint8_t>& bytes) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::setvbuf(f, nullptr, _IONBF, 0);                 // one write() of the whole file, no stdio copy
    const bool ok = std::fwrite(bytes.data(), 1, bytes.size(), f) == bytes.size();
    return std::fclose(f) == 0 && ok;
}
inline bool write_file(const std::string& path, const uint8_t* pixels, int w, int h, int channels) {
    return write_bytes(path, encode(pixels, w, h, channels));
}

// ---------------------------------------------------------------- reader
// Enough PNG to read back what this library and cv::imwrite produce: 8-bit grey / grey+alpha / RGB / RGBA,
// non-interlaced, any zlib block type (stored, fixed, dynamic Huffman), all five row filters.
// (The reference reads its own intermediates with cv::imread: src/process.cpp:217, src/mask2polygon.cpp:117,166.)
struct Image {
    int w = 0, h = 0, channels = 0;
    std::vector<uint8_t> pixels;   // row-major, `channels` bytes per pixel (RGB order for colour)
};

namespace detail {
struct BitReader {
    const uint8_t* p;
    size_t n, pos = 0;
    uint32_t buf = 0;
    int cnt = 0;
    bool ok = true;
    uint32_t bits(int k) {
        while (cnt < k) {
            if (pos >= n) { ok = false; return 0; }
            buf |= (uint32_t)p[pos++] << cnt;
            cnt += 8;
        }
        const uint32_t v = buf & ((k == 32) ? 0xFFFFFFFFu : ((1u << k) - 1u));
        buf >>= k;
        cnt -= k;
        return v;
    }
    void align() { buf = 0; cnt = 0; }
};
struct Huff {
    uint16_t count[16] = {0};
    uint16_t symbol[320] = {0};
    void build(const uint8_t* len, int n) {
        for (int i = 0; i < 16; ++i) count[i] = 0;
        for (int i = 0; i < n; ++i) count[len[i]]++;
        count[0] = 0;
        uint16_t offs[16];
        offs[1] = 0;
        for (int i = 1; i < 15; ++i) offs[i + 1] = offs[i] + count[i];
        for (int i = 0; i < n; ++i)
            if (len[i]) symbol[offs[len[i]]++] = (uint16_t)i;
    }
    int decode(BitReader& br) const {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l <= 15; ++l) {
            code |= (int)br.bits(1);
            if (!br.ok) return -1;
            const int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
        }
        return -1;
    }
};
inline bool inflate(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint16_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint16_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    if (n < 6) return false;
    BitReader br{src + 2, n - 2};   // skip the zlib header (CMF, FLG)
    int last;
    do {
        last = (int)br.bits(1);
        const int type = (int)br.bits(2);
        if (!br.ok) return false;
        if (type == 0) {
            br.align();
            if (br.pos + 4 > br.n) return false;
            const size_t len = br.p[br.pos] | (br.p[br.pos + 1] << 8);
            br.pos += 4;
            if (br.pos + len > br.n) return false;
            out.insert(out.end(), br.p + br.pos, br.p + br.pos + len);
            br.pos += len;
        } else if (type == 1 || type == 2) {
            Huff lit, dist;
            uint8_t lens[320];
            if (type == 1) {
                for (int i = 0; i < 144; ++i) lens[i] = 8;
                for (int i = 144; i < 256; ++i) lens[i] = 9;
                for (int i = 256; i < 280; ++i) lens[i] = 7;
                for (int i = 280; i < 288; ++i) lens[i] = 8;
                lit.build(lens, 288);
                for (int i = 0; i < 30; ++i) lens[i] = 5;
                dist.build(lens, 30);
            } else {
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
                if (!br.ok || nlen > 286 || ndist > 30) return false;
                uint8_t cl[19] = {0};
                for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)br.bits(3);
                Huff ch;
                ch.build(cl, 19);
                int idx = 0;
                while (idx < nlen + ndist) {
                    const int sym = ch.decode(br);
                    if (sym < 0) return false;
                    if (sym < 16) lens[idx++] = (uint8_t)sym;
                    else {
                        int rep, val = 0;
                        if (sym == 16) { if (idx == 0) return false; val = lens[idx - 1]; rep = 3 + (int)br.bits(2); }
                        else if (sym == 17) rep = 3 + (int)br.bits(3);
                        else rep = 11 + (int)br.bits(7);
                        if (idx + rep > nlen + ndist) return false;
                        while (rep--) lens[idx++] = (uint8_t)val;
                    }
                }
                lit.build(lens, nlen);
                dist.build(lens + nlen, ndist);
            }
            for (;;) {
                const int sym = lit.decode(br);
                if (sym < 0 || !br.ok) return false;
                if (sym < 256) out.push_back((uint8_t)sym);
                else if (sym == 256) break;
                else {
                    const int li = sym - 257;
                    if (li >= 29) return false;
                    const int len = lbase[li] + (int)br.bits(lext[li]);
                    const int ds = dist.decode(br);
                    if (ds < 0 || ds >= 30) return false;
                    const size_t d = dbase[ds] + br.bits(dext[ds]);
                    if (d > out.size()) return false;
                    for (int i = 0; i < len; ++i) out.push_back(out[out.size() - d]);
                }
            }
        } else {
            return false;
        }
    } while (!last);
    return true;
}
}  // namespace detail

inline bool decode(const std::vector<uint8_t>& file, Image& img) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) return false;
    size_t pos = 8;
    std::vector<uint8_t> z;
    int depth = 0, ctype = 0, interlace = 0;
    while (pos + 8 <= file.size()) {
        const uint32_t len = ((uint32_t)file[pos] << 24) | (file[pos + 1] << 16) | (file[pos + 2] << 8) | file[pos + 3];
        const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
        if (pos + 12 + (size_t)len > file.size()) return false;
        const uint8_t* d = &file[pos + 8];
        if (std::memcmp(type, "IHDR", 4) == 0 && len >= 13) {
            img.w = (int)(((uint32_t)d[0] << 24) | (d[1] << 16) | (d[2] << 8) | d[3]);
            img.h = (int)(((uint32_t)d[4] << 24) | (d[5] << 16) | (d[6] << 8) | d[7]);
            depth = d[8]; ctype = d[9]; interlace = d[12];
        } else if (std::memcmp(type, "IDAT", 4) == 0) {
            z.insert(z.end(), d, d + len);
        } else if (std::memcmp(type, "IEND", 4) == 0) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (depth != 8 || interlace != 0 || img.w <= 0 || img.h <= 0) return false;
    img.channels = ctype == 0 ? 1 : (ctype == 4 ? 2 : (ctype == 2 ? 3 : (ctype == 6 ? 4 : 0)));
    if (!img.channels) return false;
    std::vector<uint8_t> raw;
    raw.reserve(((size_t)img.w * img.channels + 1) * img.h);
    if (!detail::inflate(z.data(), z.size(), raw)) return false;
    const size_t bpp = (size_t)img.channels, row = (size_t)img.w * bpp;
    if (raw.size() < (row + 1) * (size_t)img.h) return false;
    img.pixels.assign(row * img.h, 0);
    for (int y = 0; y < img.h; ++y) {
        const uint8_t* s = &raw[(row + 1) * y];
        uint8_t* o = &img.pixels[row * y];
        const uint8_t* up = y ? o - row : nullptr;
        const int f = s[0];
        for (size_t i = 0; i < row; ++i) {
            const int a = i >= bpp ? o[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
            int pred = 0;
            if (f == 1) pred = a;
            else if (f == 2) pred = b;
            else if (f == 3) pred = (a + b) >> 1;
            else if (f == 4) {
                const int pp = a + b - c, pa = std::abs(pp - a), pb = std::abs(pp - b), pc = std::abs(pp - c);
                pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
            } else if (f != 0) return false;
            o[i] = (uint8_t)(s[1 + i] + pred);
        }
    }
    return true;
}

inline bool read_file(const std::string& path, Image& img) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::vector<uint8_t> bytes;
    uint8_t buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) bytes.insert(bytes.end(), buf, buf + n);
    std::fclose(f);
    return decode(bytes, img);
}

// grey view of any decoded image (cv::IMREAD_GRAYSCALE semantics for colour are not needed here: the masks and the
// normalised images the reference re-reads are single channel; RGB inputs are averaged the BT.601 way cv uses)
inline std::vector<uint8_t> to_grey(const Image& img) {
    std::vector<uint8_t> g((size_t)img.w * img.h);
    for (size_t i = 0; i < g.size(); ++i) {
        const uint8_t* p = &img.pixels[i * img.channels];
        g[i] = img.channels <= 2 ? p[0] : (uint8_t)((p[0] * 4899 + p[1] * 9617 + p[2] * 1868 + 8192) >> 14);
    }
    return g;
}

}  // namespace png
}  // namespace ms
