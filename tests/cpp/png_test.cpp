// Writes a synthetic frame with the headless driver's PNG writer; the Python test decodes it with zlib.
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "png_write.hpp"

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const int w = atoi(argv[2]), h = atoi(argv[3]);
    std::vector<uint32_t> rgba(size_t(w)*size_t(h));
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            rgba[size_t(y)*w + x] = 0xFF000000u | uint32_t((x*7 + y*3) & 255) | uint32_t((x ^ y) & 255) << 8 | uint32_t((x*y) & 255) << 16;
    return svo_png::writeRgb(argv[1], rgba.data(), w, h) ? 0 : 1;
}
