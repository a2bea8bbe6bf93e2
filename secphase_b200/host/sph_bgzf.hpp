// sph_bgzf.hpp -- BGZF (blocked gzip, SAM spec 4.1) stream reader/writer over zlib.
//
// Replaces, for this path only, what the reference gets from htslib's bgzf.c through
// sam_open/sam_read1 (secphase.c:236-237,268).  Blocks are independent raw-deflate members of
// <= 64 KiB, so a producer thread reads the file, cuts it into blocks and inflates a whole
// "chunk" (tens of MB) in parallel on the worker pool; the consumer sees a sequence of large
// contiguous uncompressed buffers with head room in front for carried-over bytes.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "sph_common.hpp"

namespace sph {

struct Chunk {
    uint8_t *buf = nullptr;  // malloc'ed, cap bytes
    size_t cap = 0;
    size_t head = 0;   // inflated data starts here (bytes before it are head room)
    size_t end = 0;    // one past the last inflated byte
    bool last = false; // no chunk follows
};

class BgzfReader {
public:
    BgzfReader(WorkerPool *pool, size_t chunk_bytes, size_t head_room);
    ~BgzfReader();
    int open(const char *path);
    // Next inflated chunk in stream order, or nullptr at end of stream / on error (check error()).
    Chunk *next();
    void release(Chunk *c);  // hands the buffer back to the producer
    // Makes sure `c` has at least `need` bytes of head room in front of c->head (re-allocates).
    static int grow_head(Chunk *c, size_t need);
    int error() const { return err_; }
    const std::string &error_text() const { return err_text_; }

private:
    void producer();
    bool fill(Chunk *c);
    WorkerPool *pool_;
    size_t chunk_bytes_, head_room_;
    int fd_ = -1;
    std::vector<uint8_t> comp_;  // compressed bytes not yet consumed
    size_t comp_len_ = 0;
    bool file_eof_ = false;
    std::thread th_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<Chunk *> free_, ready_;
    std::vector<Chunk *> all_;
    bool done_ = false, stop_ = false;
    int err_ = 0;
    std::string err_text_;
};

// Raw-deflate decoder for one whole BGZF block (sph_inflate.cpp): true iff exactly out_len bytes were
// produced from a well-formed stream.  `in` must have 8 readable bytes after in_len.  Callers verify the
// CRC32 and fall back to zlib on false.
bool fast_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);

// Buffers written bytes and emits them as BGZF blocks of <= 0xff00 payload bytes, deflated in
// parallel on the pool; close() appends the 28-byte end-of-file marker.
class BgzfWriter {
public:
    ~BgzfWriter();
    int open(const char *path, int level, WorkerPool *pool);
    int write(const void *data, size_t len);
    int close();

private:
    static constexpr size_t FLUSH_BYTES = 64u << 20;
    int flush(bool all);
    FILE *fp_ = nullptr;
    std::string path_;
    int level_ = 1;
    WorkerPool *pool_ = nullptr;
    std::vector<uint8_t> pend_;
};

}  // namespace sph
