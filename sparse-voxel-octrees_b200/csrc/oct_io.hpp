// .oct reader / writer (see oct_io.cpp). Host only.
#pragma once

#include <cstdint>
#include <string>

namespace svo {

struct OctFile {
    uint32_t *words = nullptr; // malloc'ed, nWords + 1 words (one zero padding word)
    uint64_t nWords = 0;
    float center[3] = {0.0f, 0.0f, 0.0f};
};

// status: 0 ok, else the svo_status value (2 io, 3 format, 4 out of memory)
bool readOctFile(const char *path, OctFile &out, std::string &err, int &status);
bool writeOctFile(const char *path, const uint32_t *words, uint64_t nWords, const float center[3], bool compress,
                  std::string &err, int &status);

} // namespace svo
