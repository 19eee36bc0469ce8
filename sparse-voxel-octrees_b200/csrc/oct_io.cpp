// .oct files: the reference's on-disk octree format, read and written byte-compatibly.
//
// Replaces VoxelOctree::VoxelOctree(const char*) (reference src/VoxelOctree.cpp:57-90)
// and VoxelOctree::save (reference src/VoxelOctree.cpp:92-123). Layout:
//   float32 center[3] | uint64 wordCount | per 64 MiB slice: uint64 compSize | LZ4 block
// where all blocks belong to ONE LZ4 streaming context: a match may reach up to
// 64 KiB back into the previous slice's plaintext, so slices decode in order
// into one contiguous buffer.
//
// The LZ4 *block format* is implemented here from its specification (token =
// literal length nibble | match length - 4 nibble, 255-saturating length
// extension bytes, literals, little-endian 16-bit offset). The reference links
// LZ4 v1.7.1 (src/third-party/lz4.h:50-52) and reads with the output-size-driven
// decoder LZ4_decompress_fast_continue, which requires of every block: the last
// 5 bytes are literals, the last match starts at least 12 bytes before the
// block end (src/third-party/lz4.c:219-223, 1172-1189). The encoder below obeys
// both, so any build of the reference loads what this file writes.
//
// Differences from the reference, on purpose: every I/O and format error is
// reported (the reference ignores fopen/fread failures, VoxelOctree.cpp:60,95),
// and slice sizes are computed in 64 bits (the reference truncates the remaining
// byte count to int, VoxelOctree.cpp:79, so it cannot load trees >= 2 GiB).
#include "oct_io.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

namespace svo {

namespace {

constexpr uint64_t kSliceBytes = 64ull*1024*1024; // CompressionBlockSize, VoxelOctree.cpp:55
constexpr size_t kMinMatch = 4;
constexpr size_t kLastLiterals = 5;
constexpr size_t kMfLimit = 12;
constexpr size_t kMaxOffset = 65535;

struct FileCloser {
    void operator()(FILE *f) const { if (f) fclose(f); }
};
typedef std::unique_ptr<FILE, FileCloser> FilePtr;

inline uint32_t read32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

// Decodes one LZ4 block into out[outPos, outPos + outSize); matches may reach
// back to out[0]. Bounded on both the input and the output side.
bool lz4DecodeBlock(const uint8_t *src, size_t srcSize, uint8_t *out, uint64_t outPos, uint64_t outSize, std::string &err) {
    const uint8_t *ip = src, *iend = src + srcSize;
    uint8_t *op = out + outPos, *oend = op + outSize;
    for (;;) {
        if (ip >= iend) { err = "LZ4 block truncated (token)"; return false; }
        unsigned token = *ip++;
        size_t litLen = token >> 4;
        if (litLen == 15) {
            unsigned s;
            do {
                if (ip >= iend) { err = "LZ4 block truncated (literal length)"; return false; }
                s = *ip++;
                litLen += s;
            } while (s == 255);
        }
        if (litLen > size_t(iend - ip) || litLen > size_t(oend - op)) { err = "LZ4 literal run overruns the block"; return false; }
        memcpy(op, ip, litLen);
        ip += litLen;
        op += litLen;
        if (op == oend) {
            if (ip != iend) { err = "LZ4 block has trailing bytes"; return false; }
            return true;
        }
        if (iend - ip < 2) { err = "LZ4 block truncated (offset)"; return false; }
        size_t offset = size_t(ip[0]) | (size_t(ip[1]) << 8);
        ip += 2;
        if (offset == 0 || offset > uint64_t(op - out)) { err = "LZ4 match offset outside the decoded data"; return false; }
        size_t matchLen = token & 15;
        if (matchLen == 15) {
            unsigned s;
            do {
                if (ip >= iend) { err = "LZ4 block truncated (match length)"; return false; }
                s = *ip++;
                matchLen += s;
            } while (s == 255);
        }
        matchLen += kMinMatch;
        if (matchLen > size_t(oend - op)) { err = "LZ4 match overruns the slice"; return false; }
        const uint8_t *match = op - offset;
        if (offset >= matchLen) {
            memcpy(op, match, matchLen);
            op += matchLen;
        } else {
            for (size_t k = 0; k < matchLen; ++k) op[k] = match[k]; // overlapping: byte order matters
            op += matchLen;
        }
    }
}

inline void putLength(std::vector<uint8_t> &out, size_t len) {
    while (len >= 255) { out.push_back(255); len -= 255; }
    out.push_back(uint8_t(len));
}

void emitSequence(std::vector<uint8_t> &out, const uint8_t *lit, size_t litLen, size_t offset, size_t matchLen) {
    size_t ml = matchLen - kMinMatch;
    out.push_back(uint8_t((litLen >= 15 ? 15 : litLen) << 4 | (ml >= 15 ? 15 : ml)));
    if (litLen >= 15) putLength(out, litLen - 15);
    out.insert(out.end(), lit, lit + litLen);
    out.push_back(uint8_t(offset & 255));
    out.push_back(uint8_t(offset >> 8));
    if (ml >= 15) putLength(out, ml - 15);
}

void emitLastLiterals(std::vector<uint8_t> &out, const uint8_t *lit, size_t litLen) {
    out.push_back(uint8_t((litLen >= 15 ? 15 : litLen) << 4));
    if (litLen >= 15) putLength(out, litLen - 15);
    out.insert(out.end(), lit, lit + litLen);
}

// Greedy single-pass LZ4 block encoder over data[begin, end) with a 64 KiB
// window that may extend into data[.., begin) (streaming context).
class Lz4Encoder {
    static constexpr int kHashBits = 16;
    std::vector<uint64_t> table_; // position + 1 of the last occurrence; 0 = none
    static inline uint32_t hash(uint32_t v) { return (v*2654435761u) >> (32 - kHashBits); }

public:
    Lz4Encoder() : table_(size_t(1) << kHashBits, 0) {}

    void encodeSlice(const uint8_t *data, uint64_t begin, uint64_t end, bool compress, std::vector<uint8_t> &out) {
        out.clear();
        const uint64_t size = end - begin;
        if (!compress || size < kMfLimit + 1) {
            emitLastLiterals(out, data + begin, size_t(size));
            return;
        }
        // entries older than the window are filtered by the offset check
        const uint64_t matchStartLimit = end - kMfLimit;   // last position a match may start at
        const uint64_t matchEndLimit = end - kLastLiterals;
        uint64_t anchor = begin, ip = begin;
        unsigned misses = 0;
        while (ip <= matchStartLimit) {
            uint32_t seq = read32(data + ip);
            uint32_t h = hash(seq);
            uint64_t cand = table_[h];
            table_[h] = ip + 1;
            if (cand != 0 && ip - (cand - 1) <= kMaxOffset && read32(data + cand - 1) == seq) {
                uint64_t m = cand - 1;
                uint64_t len = kMinMatch;
                while (ip + len < matchEndLimit && data[m + len] == data[ip + len]) ++len;
                emitSequence(out, data + anchor, size_t(ip - anchor), size_t(ip - m), size_t(len));
                // index a position inside the match so that long runs keep finding themselves
                if (ip + len - 2 <= matchStartLimit) table_[hash(read32(data + ip + len - 2))] = ip + len - 2 + 1;
                ip += len;
                anchor = ip;
                misses = 0;
            } else {
                ip += 1 + (misses++ >> 6);
            }
        }
        emitLastLiterals(out, data + anchor, size_t(end - anchor));
    }
};

bool readExact(FILE *fp, void *dst, size_t bytes) { return bytes == 0 || fread(dst, 1, bytes, fp) == bytes; }
bool writeExact(FILE *fp, const void *src, size_t bytes) { return bytes == 0 || fwrite(src, 1, bytes, fp) == bytes; }

} // namespace

bool readOctFile(const char *path, OctFile &out, std::string &err, int &status) {
    status = 0;
    FilePtr fp(fopen(path, "rb"));
    if (!fp) { err = std::string("cannot open ") + path; status = 2; return false; }
    uint64_t nWords = 0;
    if (!readExact(fp.get(), out.center, sizeof(float)*3) || !readExact(fp.get(), &nWords, 8)) {
        err = std::string("short read in the header of ") + path; status = 3; return false;
    }
    // plausibility: the rest of the file must be able to hold nWords*4 bytes at
    // LZ4's best ratio (~255:1) -- guards against garbage headers before allocating
    long here = ftell(fp.get());
    fseek(fp.get(), 0, SEEK_END);
    uint64_t fileBytes = uint64_t(ftello(fp.get()));
    fseek(fp.get(), here, SEEK_SET);
    const uint64_t totalBytes = nWords*4;
    if (nWords > (uint64_t(1) << 46) || totalBytes/256 > fileBytes) {
        err = std::string("implausible word count in ") + path; status = 3; return false;
    }
    // one padding word: the traversal kernel reads word[p + 1] next to every descriptor
    uint32_t *words = static_cast<uint32_t *>(malloc(size_t(totalBytes) + 8));
    if (!words) { err = "out of host memory for the node array"; status = 4; return false; }
    std::unique_ptr<uint32_t, void (*)(void *)> guard(words, free);
    words[nWords] = 0;
    if (nWords > 0) words[nWords - 1] = 0;

    std::vector<uint8_t> comp;
    for (uint64_t offset = 0; offset < totalBytes; offset += kSliceBytes) {
        uint64_t compSize = 0;
        if (!readExact(fp.get(), &compSize, 8)) { err = "short read (slice header)"; status = 3; return false; }
        if (compSize > fileBytes) { err = "slice larger than the file"; status = 3; return false; }
        comp.resize(size_t(compSize));
        if (!readExact(fp.get(), comp.data(), size_t(compSize))) { err = "short read (slice payload)"; status = 3; return false; }
        uint64_t outSize = totalBytes - offset < kSliceBytes ? totalBytes - offset : kSliceBytes;
        std::string lzErr;
        if (!lz4DecodeBlock(comp.data(), comp.size(), reinterpret_cast<uint8_t *>(words), offset, outSize, lzErr)) {
            err = lzErr + " in " + path; status = 3; return false;
        }
    }
    out.words = guard.release();
    out.nWords = nWords;
    return true;
}

bool writeOctFile(const char *path, const uint32_t *words, uint64_t nWords, const float center[3], bool compress,
                  std::string &err, int &status) {
    status = 0;
    FilePtr fp(fopen(path, "wb"));
    if (!fp) { err = std::string("cannot create ") + path; status = 2; return false; }
    if (!writeExact(fp.get(), center, sizeof(float)*3) || !writeExact(fp.get(), &nWords, 8)) {
        err = "write failed (header)"; status = 2; return false;
    }
    const uint8_t *data = reinterpret_cast<const uint8_t *>(words);
    const uint64_t totalBytes = nWords*4;
    Lz4Encoder enc;
    std::vector<uint8_t> comp;
    comp.reserve(size_t(kSliceBytes < totalBytes ? kSliceBytes : totalBytes) + 1024);
    for (uint64_t offset = 0; offset < totalBytes; offset += kSliceBytes) {
        uint64_t end = totalBytes - offset < kSliceBytes ? totalBytes : offset + kSliceBytes;
        enc.encodeSlice(data, offset, end, compress, comp);
        uint64_t compSize = comp.size();
        if (!writeExact(fp.get(), &compSize, 8) || !writeExact(fp.get(), comp.data(), comp.size())) {
            err = "write failed (slice)"; status = 2; return false;
        }
    }
    if (fflush(fp.get()) != 0) { err = "write failed (flush)"; status = 2; return false; }
    return true;
}

} // namespace svo
