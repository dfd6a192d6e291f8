// Host check of csrc/png_min.hpp: the SIMD checksums (PCLMULQDQ CRC-32, SSSE3 Adler-32) against the table / scalar forms on
// random buffers of awkward lengths and alignments, known answers, and encode -> decode round trips of the row-streaming
// encoder.  Prints "ok <n checks>" or the first mismatch.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../unet-medical-image-contour-segmentation-cpp_b200/csrc/png_min.hpp"

using namespace ms::png;

int main() {
    std::mt19937 rng(12345);
    int checks = 0;
    // known answers: CRC-32("123456789") = CBF43926, Adler-32("Wikipedia") = 11E60398
    if (crc32_update(0, (const uint8_t*)"123456789", 9) != 0xCBF43926u) { std::printf("crc kat\n"); return 1; }
    if (adler32((const uint8_t*)"Wikipedia", 9) != 0x11E60398u) { std::printf("adler kat\n"); return 1; }
    std::vector<uint8_t> buf(1 << 20);
    for (auto& b : buf) b = (uint8_t)rng();
    const size_t lens[] = {0, 1, 15, 16, 31, 32, 63, 64, 65, 79, 80, 127, 128, 255, 513, 1537, 5535, 5536, 5537, 5552, 11103, 65535, 65536, 262657, 788481, 1000003};
    for (size_t n : lens)
        for (size_t off : {0u, 1u, 3u, 8u, 13u}) {
            if (off + n > buf.size()) continue;
            const uint8_t* p = buf.data() + off;
            const uint32_t seed = (uint32_t)rng();
            if (crc32_update(seed, p, n) != (n ? crc32_table(seed, p, n) : seed)) { std::printf("crc mismatch n=%zu off=%zu\n", n, off); return 1; }
            Adler a, b;
            a.a = b.a = 1 + rng() % 65520; a.b = b.b = rng() % 65521;
            adler_update(a, p, n);
            adler_scalar(b, p, n);
            if (a.a != b.a || a.b != b.b) { std::printf("adler mismatch n=%zu off=%zu\n", n, off); return 1; }
            checks += 2;
        }
    // worst case for the 32-bit lanes: all 0xFF
    std::vector<uint8_t> ff(200000, 0xFF);
    Adler a, b;
    adler_update(a, ff.data(), ff.size());
    adler_scalar(b, ff.data(), ff.size());
    if (a.value() != b.value()) { std::printf("adler 0xFF\n"); return 1; }
    // incremental == one shot (rows are fed one at a time by the encoder)
    Adler inc;
    for (size_t i = 0; i < 100000; i += 513) adler_update(inc, buf.data() + i, std::min<size_t>(513, 100000 - i));
    if (inc.value() != adler32(buf.data(), 100000)) { std::printf("adler incremental\n"); return 1; }
    // encode -> decode round trips: sizes around the 65,535-byte stored-block boundary, grey and RGB
    const int shapes[][3] = {{512, 512, 1}, {512, 512, 3}, {1, 1, 1}, {7, 3, 3}, {65534, 1, 1}, {65535, 1, 1}, {65536, 1, 1}, {333, 517, 3}, {4, 0, 1}};
    for (const auto& s : shapes) {
        const int w = s[0], h = s[1], c = s[2];
        std::vector<uint8_t> px((size_t)w * h * c);
        for (auto& v : px) v = (uint8_t)rng();
        const std::vector<uint8_t> file = encode(px.data(), w, h, c);
        if (h == 0) { ++checks; continue; }
        Image img;
        if (!decode(file, img) || img.w != w || img.h != h || img.channels != c || img.pixels != px) { std::printf("png round trip %dx%dx%d\n", w, h, c); return 1; }
        // every chunk CRC and the zlib Adler are checked by re-deriving them with the table / scalar forms
        size_t pos = 8;
        while (pos + 12 <= file.size()) {
            const uint32_t len = ((uint32_t)file[pos] << 24) | (file[pos + 1] << 16) | (file[pos + 2] << 8) | file[pos + 3];
            const uint32_t want = ((uint32_t)file[pos + 8 + len] << 24) | (file[pos + 9 + len] << 16) | (file[pos + 10 + len] << 8) | file[pos + 11 + len];
            if (crc32_table(0, &file[pos + 4], 4 + len) != want) { std::printf("chunk crc %dx%dx%d\n", w, h, c); return 1; }
            pos += 12 + len;
        }
        if (pos != file.size()) { std::printf("trailing bytes\n"); return 1; }
        ++checks;
    }
    std::printf("ok %d\n", checks);
    return 0;
}
