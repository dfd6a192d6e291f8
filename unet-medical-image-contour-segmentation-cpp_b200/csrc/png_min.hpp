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
inline uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
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
// Adler-32 with the modulo deferred to once per 5552 bytes (the largest run that cannot overflow 32 bits)
inline uint32_t adler32(const uint8_t* p, size_t n) {
    uint32_t a = 1, b = 0;
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
    return (b << 16) | a;
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

inline std::vector<uint8_t> encode(const uint8_t* pixels, int w, int h, int channels) {
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    put32(ihdr, (uint32_t)w);
    put32(ihdr, (uint32_t)h);
    ihdr.push_back(8);
    ihdr.push_back(channels == 3 ? 2 : 0);
    ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(out, "IHDR", ihdr);
    // raw scanlines, filter type 0
    const size_t row = (size_t)w * channels, raw_n = (row + 1) * (size_t)h;
    std::vector<uint8_t> raw(raw_n);
    for (int y = 0; y < h; ++y) {
        uint8_t* dst = raw.data() + (size_t)y * (row + 1);
        dst[0] = 0;
        std::memcpy(dst + 1, pixels + (size_t)y * row, row);
    }
    // IDAT = zlib stream of stored blocks, written in place: length | "IDAT" | 78 01 | blocks | adler | crc
    const size_t n_blocks = raw_n == 0 ? 1 : (raw_n + 65534) / 65535;
    const size_t z_n = 2 + n_blocks * 5 + raw_n + 4;
    const size_t idat_at = out.size();
    out.resize(idat_at + 4 + 4 + z_n + 4);
    uint8_t* q = out.data() + idat_at;
    store32(q, (uint32_t)z_n);
    std::memcpy(q + 4, "IDAT", 4);
    uint8_t* z = q + 8;
    *z++ = 0x78;
    *z++ = 0x01;
    size_t pos = 0;
    for (size_t blk = 0; blk < n_blocks; ++blk) {
        const size_t n = std::min<size_t>(65535, raw_n - pos);
        *z++ = blk + 1 == n_blocks ? 1 : 0;
        *z++ = (uint8_t)(n & 0xFF);
        *z++ = (uint8_t)(n >> 8);
        *z++ = (uint8_t)(~n & 0xFF);
        *z++ = (uint8_t)((~n >> 8) & 0xFF);
        std::memcpy(z, raw.data() + pos, n);
        z += n;
        pos += n;
    }
    store32(z, adler32(raw.data(), raw_n));
    z += 4;
    store32(z, crc32_update(0, q + 4, 4 + z_n));
    chunk(out, "IEND", {});
    return out;
}

inline bool write_file(const std::string& path, const uint8_t* pixels, int w, int h, int channels) {
    std::vector<uint8_t> bytes = encode(pixels, w, h, channels);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(bytes.data(), 1, bytes.size(), f) == bytes.size();
    return std::fclose(f) == 0 && ok;
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
