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
// Loading is pipelined and, where the file allows it, parallel ("next" row f1 of
// SURVEY.md section 8): worker threads take slices in order and decode them
// concurrently into their own part of the node array. A slice whose matches
// reach back into the previous slice (files written by the reference's
// LZ4_compress_continue do, within the first 64 KiB of plaintext) blocks at
// that match until its predecessor is complete; slices without such matches
// (everything writeOctFile below produces) never block. Completed slices are
// handed to the caller in order while later ones are still being decoded, which
// is how svo_tree_load_oct overlaps the host->device copy with the decode.
//
// Differences from the reference, on purpose: every I/O and format error is
// reported (the reference ignores fopen/fread failures, VoxelOctree.cpp:60,95),
// and slice sizes are computed in 64 bits (the reference truncates the remaining
// byte count to int, VoxelOctree.cpp:79, so it cannot load trees >= 2 GiB).
#include "oct_io.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include <sys/mman.h>

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
// back to out[0]. Bounded on both the input and the output side. `waitPrev` is
// called once, before the first match that starts in plaintext preceding
// outPos is copied (another thread may still be producing it).
//
// Copies are "wild" where there is room: 32-byte literal and up to 16-byte
// match chunks that may write past the end of the run, but never past this
// slice's own output range, so concurrent decoders of neighbouring slices do
// not interfere. Near the end of the slice overlapping matches (offset <
// length: runs) are expanded by doubling: once `offset` bytes are in place the
// destination is periodic, so every further memcpy can be as long as
// everything written so far.
template <class WaitPrev>
bool lz4DecodeBlock(const uint8_t *src, size_t srcSize, uint8_t *out, uint64_t outPos, uint64_t outSize,
                    WaitPrev &&waitPrev, std::string &err) {
    const uint8_t *ip = src, *const iend = src + srcSize;
    uint8_t *const start = out + outPos;
    uint8_t *op = start, *const oend = op + outSize;
    bool prevReady = outPos == 0;
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
        if (litLen <= 32 && size_t(iend - ip) >= 32 && size_t(oend - op) >= 32)
            memcpy(op, ip, 32);
        else
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
        if (match < start && !prevReady) {
            if (!waitPrev()) { err = "previous slice failed"; return false; }
            prevReady = true;
        }
        if (size_t(oend - op) >= matchLen + 32) {
            // chunked copies that may run past the match (never past the slice); a chunk no longer than
            // the offset reads only bytes that are already final, overlapping match or not. The first 24
            // (16) bytes are copied without a loop: 95 % of the matches in octree data are shorter, and
            // one well-predicted offset test (92 % are >= 8) beats wider chunks behind a 16-byte test.
            uint8_t *d = op, *const e = op + matchLen;
            const uint8_t *m = match;
            if (offset >= 8) {
                memcpy(d, m, 8);
                memcpy(d + 8, m + 8, 8);
                memcpy(d + 16, m + 16, 8);
                if (matchLen > 24) {
                    d += 24; m += 24;
                    do { memcpy(d, m, 8); d += 8; m += 8; } while (d < e);
                }
            } else if (offset >= 4) {
                memcpy(d, m, 4);
                memcpy(d + 4, m + 4, 4);
                memcpy(d + 8, m + 8, 4);
                memcpy(d + 12, m + 12, 4);
                if (matchLen > 16) {
                    d += 16; m += 16;
                    do { memcpy(d, m, 4); d += 4; m += 4; } while (d < e);
                }
            } else {
                do { *d++ = *m++; } while (d < e);
            }
        } else if (offset >= matchLen) {
            memcpy(op, match, matchLen);
        } else {
            size_t done = 0;
            while (done < matchLen) {   // [match, op + done) is periodic with period `offset`; done % offset == 0
                size_t n = std::min(offset + done, matchLen - done);
                memcpy(op + done, match, n);
                done += n;
            }
        }
        op += matchLen;
    }
}

// Output cursor over a buffer sized for the worst case (an incompressible slice is one literal run:
// size + size/255 + 16 bytes; a match never costs more bytes than the 4+ it replaces).
struct ByteSink {
    uint8_t *p;
    inline void put(uint8_t v) { *p++ = v; }
    inline void length(size_t len) {
        while (len >= 255) { *p++ = 255; len -= 255; }
        *p++ = uint8_t(len);
    }
    inline void literals(const uint8_t *lit, size_t n) { memcpy(p, lit, n); p += n; }
};

inline void emitSequence(ByteSink &out, const uint8_t *lit, size_t litLen, size_t offset, size_t matchLen) {
    size_t ml = matchLen - kMinMatch;
    out.put(uint8_t((litLen >= 15 ? 15 : litLen) << 4 | (ml >= 15 ? 15 : ml)));
    if (litLen >= 15) out.length(litLen - 15);
    out.literals(lit, litLen);
    out.put(uint8_t(offset & 255));
    out.put(uint8_t(offset >> 8));
    if (ml >= 15) out.length(ml - 15);
}

inline void emitLastLiterals(ByteSink &out, const uint8_t *lit, size_t litLen) {
    out.put(uint8_t((litLen >= 15 ? 15 : litLen) << 4));
    if (litLen >= 15) out.length(litLen - 15);
    out.literals(lit, litLen);
}

// Greedy single-pass LZ4 block encoder over data[begin, end) with a 64 KiB
// window. The window never extends into data[.., begin): still a valid stream
// for the reference's LZ4_decompress_fast_continue, which merely allows such
// matches.
class Lz4Encoder {
    static constexpr int kHashBits = 16;
    std::vector<uint32_t> table_; // (position - begin) + 1 of the last occurrence; 0 = none (slices are <= 64 MiB)
    // Hash of FIVE bytes (the candidate test below still compares four): octree words repeat so often that a table
    // keyed on four bytes keeps replacing a position by a later one with the same word and a shorter match behind
    // it. Measured on a 64 MiB slice of the 8192^3 tree: 32.9 -> 30.0 MB (the reference's LZ4 1.7.1: 31.2 MB).
    static inline uint32_t hash(const uint8_t *p) {
        uint64_t v;
        memcpy(&v, p, 8);                           // callers stay 12 bytes clear of the slice end
        return uint32_t(((v << 24)*889523592379ull) >> (64 - kHashBits));
    }

public:
    Lz4Encoder() : table_(size_t(1) << kHashBits, 0) {}

    void encodeSlice(const uint8_t *data, uint64_t begin, uint64_t end, bool compress, std::vector<uint8_t> &buf) {
        const uint64_t size = end - begin;
        buf.resize(size_t(size + size/255 + 64));
        ByteSink out{buf.data()};
        if (!compress || size < kMfLimit + 1) {
            emitLastLiterals(out, data + begin, size_t(size));
            buf.resize(size_t(out.p - buf.data()));
            return;
        }
        const uint8_t *base = data + begin;                 // positions below are relative to the slice
        const uint64_t matchStartLimit = size - kMfLimit;   // last position a match may start at
        const uint64_t matchEndLimit = size - kLastLiterals;
        uint64_t anchor = 0, ip = 0;
        unsigned misses = 0;
        while (ip <= matchStartLimit) {
            const uint32_t seq = read32(base + ip);
            const uint32_t h = hash(base + ip);
            const uint32_t cand = table_[h];
            table_[h] = uint32_t(ip) + 1;
            // the table only ever holds positions of this slice: slices this writer produces never refer to
            // each other's plaintext, so the reader can decode them concurrently
            if (cand != 0 && ip - (cand - 1) <= kMaxOffset && read32(base + cand - 1) == seq) {
                uint64_t m = cand - 1;
                // catch up: the match may have begun before the position the table happened to hold (-2 % bytes, +35 % speed)
                while (ip > anchor && m > 0 && base[ip - 1] == base[m - 1]) { --ip; --m; }
                uint64_t len = kMinMatch;
                while (ip + len + 8 <= matchEndLimit) {     // eight bytes at a time, then the tail
                    uint64_t x, y;
                    memcpy(&x, base + m + len, 8);
                    memcpy(&y, base + ip + len, 8);
                    if (x != y) { len += uint64_t(__builtin_ctzll(x ^ y)) >> 3; goto matched; }
                    len += 8;
                }
                while (ip + len < matchEndLimit && base[m + len] == base[ip + len]) ++len;
            matched:
                emitSequence(out, base + anchor, size_t(ip - anchor), size_t(ip - m), size_t(len));
                // index a position inside the match so that long runs keep finding themselves
                if (ip + len - 2 <= matchStartLimit) table_[hash(base + ip + len - 2)] = uint32_t(ip + len - 2) + 1;
                ip += len;
                anchor = ip;
                misses = 0;
            } else {
                ip += 1 + (misses++ >> 6);
            }
        }
        emitLastLiterals(out, base + anchor, size_t(size - anchor));
        buf.resize(size_t(out.p - buf.data()));
    }
};

bool readExact(FILE *fp, void *dst, size_t bytes) { return bytes == 0 || fread(dst, 1, bytes, fp) == bytes; }
bool writeExact(FILE *fp, const void *src, size_t bytes) { return bytes == 0 || fwrite(src, 1, bytes, fp) == bytes; }

int ioThreads(uint64_t slices) {
    int want = 0;
    if (const char *e = getenv("SVO_IO_THREADS")) want = atoi(e);
    if (want <= 0) {
        want = int(std::thread::hardware_concurrency());
        if (want > 16) want = 16;
    }
    if (want < 1) want = 1;
    if (uint64_t(want) > slices) want = int(slices ? slices : 1);
    return want;
}

struct Slice {
    uint64_t fileOffset; // of the LZ4 payload
    uint64_t compSize;
    uint64_t outPos, outSize;
};

} // namespace

OctReader::~OctReader() { close(); }

void OctReader::close() {
    if (fp_) fclose(static_cast<FILE *>(fp_));
    fp_ = nullptr;
}

bool OctReader::open(const char *path, std::string &err, int &status) {
    status = 0;
    close();
    path_ = path;
    FILE *fp = fopen(path, "rb");
    if (!fp) { err = std::string("cannot open ") + path; status = 2; return false; }
    fp_ = fp;
    nWords = 0;
    if (!readExact(fp, center, sizeof(float)*3) || !readExact(fp, &nWords, 8)) {
        err = std::string("short read in the header of ") + path; status = 3; return false;
    }
    // plausibility: the rest of the file must be able to hold nWords*4 bytes at
    // LZ4's best ratio (~255:1) -- guards against garbage headers before allocating
    off_t here = ftello(fp);
    fseeko(fp, 0, SEEK_END);
    fileBytes_ = uint64_t(ftello(fp));
    fseeko(fp, here, SEEK_SET);
    if (nWords > (uint64_t(1) << 46) || nWords*4/256 > fileBytes_) {
        err = std::string("implausible word count in ") + path; status = 3; return false;
    }
    return true;
}

bool OctReader::decode(uint32_t *words, const SliceSink &sink, std::string &err, int &status) {
    status = 0;
    FILE *fp = static_cast<FILE *>(fp_);
    if (!fp) { err = "OctReader::decode without open"; status = 1; return false; }
    const uint64_t totalBytes = nWords*4;

    // slice table: the compressed sizes chain through the file
    std::vector<Slice> slices;
    uint64_t pos = uint64_t(ftello(fp));
    for (uint64_t offset = 0; offset < totalBytes; offset += kSliceBytes) {
        uint64_t compSize = 0;
        if (fseeko(fp, off_t(pos), SEEK_SET) != 0 || !readExact(fp, &compSize, 8)) { err = "short read (slice header)"; status = 3; return false; }
        if (compSize > fileBytes_ || pos + 8 + compSize > fileBytes_) { err = "short read (slice payload): slice larger than the file"; status = 3; return false; }
        slices.push_back({pos + 8, compSize, offset, std::min(kSliceBytes, totalBytes - offset)});
        pos += 8 + compSize;
    }
    const size_t n = slices.size();
    if (n == 0) return true;

    // state shared by the workers and the delivering thread
    std::mutex mu;
    std::condition_variable cv;
    std::vector<char> done(n, 0);
    bool failed = false;
    std::string firstErr;
    std::atomic<size_t> next(0);
    uint8_t *out = reinterpret_cast<uint8_t *>(words);
    const std::string path = path_;

    auto fail = [&](const std::string &what) {
        std::lock_guard<std::mutex> lock(mu);
        if (!failed) { failed = true; firstErr = what; }
        cv.notify_all();
    };
    auto worker = [&]() {
        FilePtr f(fopen(path.c_str(), "rb"));   // own handle: own file position
        std::vector<uint8_t> comp;
        if (!f) { fail("cannot reopen " + path); return; }
        try {
        for (;;) {
            size_t k = next.fetch_add(1);
            if (k >= n) return;
            {
                std::lock_guard<std::mutex> lock(mu);
                if (failed) return;
            }
            const Slice &sl = slices[k];
            // Fault this slice's pages in now, while the predecessor may still be decoding: first-touch
            // page faults cost more than the decode itself on a fresh 64 MiB range.
            for (uint64_t b = 0; b < sl.outSize; b += 4096) out[sl.outPos + b] = 0;
            comp.resize(size_t(sl.compSize));
            if (fseeko(f.get(), off_t(sl.fileOffset), SEEK_SET) != 0 || !readExact(f.get(), comp.data(), comp.size())) {
                fail("short read (slice payload)");
                return;
            }
            std::string lzErr;
            auto waitPrev = [&]() {
                std::unique_lock<std::mutex> lock(mu);
                cv.wait(lock, [&] { return failed || done[k - 1]; });
                return !failed;
            };
            if (!lz4DecodeBlock(comp.data(), comp.size(), out, sl.outPos, sl.outSize, waitPrev, lzErr)) {
                fail(lzErr);
                return;
            }
            {
                std::lock_guard<std::mutex> lock(mu);
                done[k] = 1;
            }
            cv.notify_all();
        }
        } catch (const std::exception &e) {     // e.g. bad_alloc for a slice payload: report, do not terminate
            fail(std::string("slice decoder: ") + e.what());
        }
    };

    const int nThreads = ioThreads(n);
    std::vector<std::thread> pool;
    for (int t = 0; t < nThreads; ++t) pool.emplace_back(worker);

    // deliver finished slices in order while the rest is being decoded
    bool ok = true;
    for (size_t k = 0; k < n && ok; ++k) {
        {
            std::unique_lock<std::mutex> lock(mu);
            cv.wait(lock, [&] { return failed || done[k]; });
            if (failed) { ok = false; break; }
        }
        if (sink && !sink(slices[k].outPos/4, slices[k].outSize/4)) {
            fail("slice consumer failed");
            ok = false;
        }
    }
    for (auto &t : pool) t.join();
    if (failed) {
        err = firstErr + " in " + path_;
        status = 3;
        return false;
    }
    return ok;
}

// 2 MiB alignment + MADV_HUGEPAGE: with transparent huge pages the 1.6 GB of an 8192^3 tree is 800
// page faults instead of 400,000. free()-compatible (svo_free releases it). Room for one padding word.
uint32_t *allocNodeArray(uint64_t nWords) {
    const size_t bytes = size_t(nWords)*4 + 8;
    void *mem = nullptr;
    if (posix_memalign(&mem, size_t(2) << 20, bytes) != 0) return nullptr;
#ifdef MADV_HUGEPAGE
    madvise(mem, bytes, MADV_HUGEPAGE);
#endif
    return static_cast<uint32_t *>(mem);
}

bool readOctFile(const char *path, OctFile &out, std::string &err, int &status) {
    OctReader reader;
    if (!reader.open(path, err, status)) return false;
    // one padding word: the traversal kernel reads word[p + 1] next to every descriptor
    uint32_t *words = allocNodeArray(reader.nWords);
    if (!words) { err = "out of host memory for the node array"; status = 4; return false; }
    std::unique_ptr<uint32_t, void (*)(void *)> guard(words, free);
    words[reader.nWords] = 0;
    if (!reader.decode(words, nullptr, err, status)) return false;
    out.words = guard.release();
    out.nWords = reader.nWords;
    memcpy(out.center, reader.center, sizeof(float)*3);
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
    const uint64_t nSlices = (totalBytes + kSliceBytes - 1)/kSliceBytes;
    // slices are independent, so a batch of them is encoded concurrently and written in order
    const int nThreads = ioThreads(nSlices);
    std::vector<std::vector<uint8_t>> comp(static_cast<size_t>(nThreads));
    for (uint64_t first = 0; first < nSlices; first += uint64_t(nThreads)) {
        const int batch = int(std::min<uint64_t>(uint64_t(nThreads), nSlices - first));
        auto encode = [&](int j) {
            Lz4Encoder enc;
            uint64_t begin = (first + uint64_t(j))*kSliceBytes;
            uint64_t end = std::min(begin + kSliceBytes, totalBytes);
            enc.encodeSlice(data, begin, end, compress, comp[size_t(j)]);
        };
        std::vector<std::thread> pool;
        for (int j = 1; j < batch; ++j) pool.emplace_back(encode, j);
        encode(0);
        for (auto &t : pool) t.join();
        for (int j = 0; j < batch; ++j) {
            uint64_t compSize = comp[size_t(j)].size();
            if (!writeExact(fp.get(), &compSize, 8) || !writeExact(fp.get(), comp[size_t(j)].data(), comp[size_t(j)].size())) {
                err = "write failed (slice)"; status = 2; return false;
            }
        }
    }
    if (fflush(fp.get()) != 0) { err = "write failed (flush)"; status = 2; return false; }
    return true;
}

} // namespace svo
