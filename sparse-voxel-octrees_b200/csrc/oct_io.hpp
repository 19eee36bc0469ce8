// .oct reader / writer (see oct_io.cpp). Host only.
#pragma once

#include <cstdint>
#include <functional>
#include <string>

namespace svo {

struct OctFile {
    uint32_t *words = nullptr; // malloc'ed, nWords + 1 words (one zero padding word)
    uint64_t nWords = 0;
    float center[3] = {0.0f, 0.0f, 0.0f};
};

// Streaming reader: open() parses the header, decode() fills `words` (room for nWords words) with
// worker threads (SVO_IO_THREADS, default min(cores, 16)) and calls `sink(firstWord, wordCount)` from
// the calling thread for every 64 MiB slice, in file order, as soon as it is complete -- later slices
// are still being decoded meanwhile. A sink returning false aborts the load.
class OctReader {
public:
    typedef std::function<bool(uint64_t firstWord, uint64_t wordCount)> SliceSink;
    uint64_t nWords = 0;
    float center[3] = {0.0f, 0.0f, 0.0f};
    ~OctReader();
    bool open(const char *path, std::string &err, int &status);
    bool decode(uint32_t *words, const SliceSink &sink, std::string &err, int &status);
    void close();

private:
    void *fp_ = nullptr;
    uint64_t fileBytes_ = 0;
    std::string path_;
};

// Host buffer for nWords + 1 words, huge-page friendly; release with free().
uint32_t *allocNodeArray(uint64_t nWords);

// status: 0 ok, else the svo_status value (2 io, 3 format, 4 out of memory)
bool readOctFile(const char *path, OctFile &out, std::string &err, int &status);
bool writeOctFile(const char *path, const uint32_t *words, uint64_t nWords, const float center[3], bool compress,
                  std::string &err, int &status);

} // namespace svo
