// Minimal PNG writer for the headless driver (the window's frames as image files).
// 8-bit RGB, filter 0 on every row, zlib stream of *stored* deflate blocks (no compression: a frame is written at disk
// speed and any PNG reader opens it), CRC-32 per chunk and Adler-32 over the raw scanlines. Written from the PNG /
// zlib / deflate specifications; no library.
#ifndef SVO_B200_PNG_WRITE_HPP_
#define SVO_B200_PNG_WRITE_HPP_

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace svo_png {

inline uint32_t crc32(const uint8_t *p, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 255] ^ (crc >> 8);
    return ~crc;
}

inline void put32(std::vector<uint8_t> &v, uint32_t x) {
    v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x));
}

inline bool writeChunk(FILE *fp, const char type[4], const std::vector<uint8_t> &data) {
    std::vector<uint8_t> head;
    put32(head, uint32_t(data.size()));
    head.insert(head.end(), type, type + 4);
    uint32_t crc = crc32(head.data() + 4, 4);
    crc = crc32(data.data(), data.size(), crc);
    std::vector<uint8_t> tail;
    put32(tail, crc);
    return fwrite(head.data(), 1, head.size(), fp) == head.size() &&
           (data.empty() || fwrite(data.data(), 1, data.size(), fp) == data.size()) &&
           fwrite(tail.data(), 1, 4, fp) == 4;
}

// rgba: the frame as svo_render_frame returns it (0xAABBGGRR words, row-major); alpha is dropped like in the PPM path
inline bool writeRgb(const std::string &path, const uint32_t *rgba, int w, int h) {
    if (w <= 0 || h <= 0) return false;
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp) return false;
    static const uint8_t signature[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    bool ok = fwrite(signature, 1, 8, fp) == 8;
    std::vector<uint8_t> ihdr;
    put32(ihdr, uint32_t(w));
    put32(ihdr, uint32_t(h));
    const uint8_t rest[5] = {8, 2, 0, 0, 0};       // 8 bits, colour type 2 (RGB), deflate, adaptive filtering, no interlace
    ihdr.insert(ihdr.end(), rest, rest + 5);
    ok = ok && writeChunk(fp, "IHDR", ihdr);

    const size_t rowBytes = size_t(w)*3 + 1;       // filter byte + pixels
    std::vector<uint8_t> raw(rowBytes*size_t(h));
    for (int y = 0; y < h; ++y) {
        uint8_t *row = raw.data() + rowBytes*size_t(y);
        row[0] = 0;
        for (int x = 0; x < w; ++x) {
            const uint32_t p = rgba[size_t(y)*size_t(w) + size_t(x)];
            row[1 + 3*x] = uint8_t(p); row[2 + 3*x] = uint8_t(p >> 8); row[3 + 3*x] = uint8_t(p >> 16);
        }
    }
    uint32_t a = 1, b = 0;                          // Adler-32
    for (size_t i = 0; i < raw.size();) {
        const size_t n = raw.size() - i < 5552 ? raw.size() - i : 5552;
        for (size_t k = 0; k < n; ++k) { a += raw[i + k]; b += a; }
        a %= 65521; b %= 65521;
        i += n;
    }
    std::vector<uint8_t> z;
    z.reserve(raw.size() + raw.size()/65535*5 + 16);
    z.push_back(0x78); z.push_back(0x01);           // zlib header: deflate, 32 KiB window, no preset dictionary
    for (size_t i = 0; i < raw.size();) {
        const size_t n = raw.size() - i < 65535 ? raw.size() - i : 65535;
        z.push_back(i + n == raw.size() ? 1 : 0);   // BFINAL, BTYPE = 00 (stored)
        z.push_back(uint8_t(n)); z.push_back(uint8_t(n >> 8));
        z.push_back(uint8_t(~n)); z.push_back(uint8_t((~n) >> 8));
        z.insert(z.end(), raw.begin() + long(i), raw.begin() + long(i + n));
        i += n;
    }
    put32(z, (b << 16) | a);
    ok = ok && writeChunk(fp, "IDAT", z);
    ok = ok && writeChunk(fp, "IEND", std::vector<uint8_t>());
    return fclose(fp) == 0 && ok;
}

} // namespace svo_png

#endif
