// png_min.hpp -- dependency-free PNG writer for the reference's side artefacts.
//
// The reference writes `_normalized.png` and `_mask.png` with cv::imwrite(..., PNG_COMPRESSION 0)
// (/root/reference/src/preprocess.cpp:122, src/process.cpp:236-239) and the BGR overlay with default
// compression (src/mask2polygon.cpp:126).  PNG is lossless, so parity is defined on decoded pixels,
// not on file bytes; this writer emits valid PNGs with stored (uncompressed) deflate blocks.
// 8-bit grayscale (channels = 1) or 8-bit RGB (channels = 3, caller passes RGB order).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace ms {
namespace png {

inline uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}

inline void put32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
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
    const size_t row = (size_t)w * channels;
    std::vector<uint8_t> raw;
    raw.reserve((row + 1) * h);
    for (int y = 0; y < h; ++y) {
        raw.push_back(0);
        raw.insert(raw.end(), pixels + (size_t)y * row, pixels + (size_t)(y + 1) * row);
    }
    // zlib stream of stored blocks
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (uint8_t c : raw) { a = (a + c) % 65521u; b = (b + a) % 65521u; }
    size_t pos = 0;
    do {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back((uint8_t)(n & 0xFF)); z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 0xFF)); z.push_back((uint8_t)((~n >> 8) & 0xFF));
        z.insert(z.end(), raw.begin() + (long)pos, raw.begin() + (long)(pos + n));
        pos += n;
    } while (pos < raw.size());
    put32(z, (b << 16) | a);
    chunk(out, "IDAT", z);
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

}  // namespace png
}  // namespace ms
