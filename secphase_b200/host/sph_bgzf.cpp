#include "sph_bgzf.hpp"

#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cstdlib>
#include <cstring>

#include "../../include/secphase_host.h"

namespace sph {

namespace {

struct BlockRef {
    size_t comp_off;  // start of the raw deflate stream inside comp_
    uint32_t comp_len;
    uint32_t ulen;    // ISIZE
    uint32_t crc;
    size_t uoff;      // destination offset inside the chunk buffer
};

inline uint32_t le32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t) p[3] << 24); }
inline uint16_t le16(const uint8_t *p) { return (uint16_t) (p[0] | (p[1] << 8)); }

// Parses one BGZF member header at p (avail bytes).  Returns the total block size, 0 if more
// bytes are needed, -1 if this is not a BGZF block.
long bgzf_block_size(const uint8_t *p, size_t avail, size_t *payload_off) {
    if (avail < 18) return 0;
    if (p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) return -1;
    size_t xlen = le16(p + 10);
    if (avail < 12 + xlen) return 0;
    size_t o = 12, end = 12 + xlen;
    long bsize = -1;
    while (o + 4 <= end) {
        uint8_t si1 = p[o], si2 = p[o + 1];
        size_t slen = le16(p + o + 2);
        if (si1 == 'B' && si2 == 'C' && slen == 2 && o + 6 <= end) bsize = (long) le16(p + o + 4) + 1;
        o += 4 + slen;
    }
    if (bsize < 0 || (size_t) bsize < end + 8) return -1;
    *payload_off = end;
    return bsize;
}

}  // namespace

BgzfReader::BgzfReader(WorkerPool *pool, size_t chunk_bytes, size_t head_room)
    : pool_(pool), chunk_bytes_(chunk_bytes), head_room_(head_room) {}

BgzfReader::~BgzfReader() {
    {
        std::lock_guard<std::mutex> g(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
    for (Chunk *c : all_) {
        free(c->buf);
        delete c;
    }
    if (fd_ >= 0) close(fd_);
}

int BgzfReader::open(const char *path) {
    fd_ = ::open(path, O_RDONLY);
    if (fd_ < 0) {
        set_error("cannot open %s: %s", path, strerror(errno));
        return SPH_EIO;
    }
#ifdef POSIX_FADV_SEQUENTIAL
    posix_fadvise(fd_, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
    comp_.resize(chunk_bytes_ + (1u << 20));
    for (int i = 0; i < 3; i++) {
        Chunk *c = new Chunk();
        c->cap = head_room_ + chunk_bytes_ + (1u << 17);
        c->buf = (uint8_t *) malloc(c->cap);
        if (!c->buf) {
            delete c;
            set_error("out of memory for a %zu-byte BGZF chunk", head_room_ + chunk_bytes_);
            return SPH_ENOMEM;
        }
        all_.push_back(c);
        free_.push_back(c);
    }
    th_ = std::thread([this] { producer(); });
    return SPH_OK;
}

int BgzfReader::grow_head(Chunk *c, size_t need) {
    if (c->head >= need) return SPH_OK;
    size_t len = c->end - c->head;
    size_t ncap = need + len + (1u << 16);
    uint8_t *nb = (uint8_t *) malloc(ncap);
    if (!nb) return SPH_ENOMEM;
    memcpy(nb + need, c->buf + c->head, len);
    free(c->buf);
    c->buf = nb;
    c->cap = ncap;
    c->head = need;
    c->end = need + len;
    return SPH_OK;
}

// Reads compressed bytes, cuts them into blocks until ~chunk_bytes_ of output are described,
// then inflates the blocks in parallel into c.  Returns false at a clean end of stream with
// nothing produced, or on error (err_ set).
bool BgzfReader::fill(Chunk *c) {
    c->head = head_room_;
    c->end = c->head;
    c->last = false;
    if (c->cap < head_room_ + chunk_bytes_ + (1u << 17)) {  // a grown-then-shrunk buffer
        free(c->buf);
        c->cap = head_room_ + chunk_bytes_ + (1u << 17);
        c->buf = (uint8_t *) malloc(c->cap);
        if (!c->buf) {
            err_ = SPH_ENOMEM;
            err_text_ = "out of memory";
            return false;
        }
    }
    std::vector<BlockRef> blocks;
    size_t pos = 0, out = 0;
    for (;;) {
        // top up the compressed buffer
        while (!file_eof_ && comp_len_ < comp_.size()) {
            ssize_t n = ::read(fd_, comp_.data() + comp_len_, comp_.size() - comp_len_);
            if (n < 0) {
                if (errno == EINTR) continue;
                err_ = SPH_EIO;
                err_text_ = std::string("read failed: ") + strerror(errno);
                return false;
            }
            if (n == 0) file_eof_ = true;
            comp_len_ += (size_t) n;
        }
        bool need_more = false;
        while (pos < comp_len_ && out < chunk_bytes_) {
            size_t payload = 0;
            long bs = bgzf_block_size(comp_.data() + pos, comp_len_ - pos, &payload);
            if (bs < 0) {
                err_ = SPH_EFORMAT;
                err_text_ = "not a BGZF block (is the input a BAM file?)";
                return false;
            }
            if (bs == 0 || pos + (size_t) bs > comp_len_) {
                need_more = true;
                break;
            }
            const uint8_t *b = comp_.data() + pos;
            BlockRef r;
            r.comp_off = pos + payload;
            r.comp_len = (uint32_t) ((size_t) bs - payload - 8);
            r.crc = le32(b + bs - 8);
            r.ulen = le32(b + bs - 4);
            if (r.ulen > 65536) {
                err_ = SPH_EFORMAT;
                err_text_ = "BGZF block larger than 64 KiB";
                return false;
            }
            r.uoff = c->head + out;
            out += r.ulen;
            if (r.ulen) blocks.push_back(r);
            pos += (size_t) bs;
        }
        if (out >= chunk_bytes_) break;
        if (!need_more) {  // every buffered byte was consumed
            if (out > 0 || file_eof_) break;
            comp_len_ = 0;  // nothing but empty blocks so far
            pos = 0;
            continue;
        }
        if (file_eof_) {
            err_ = SPH_EFORMAT;
            err_text_ = "truncated BGZF block at end of file";
            return false;
        }
        if (pos == 0 && comp_len_ == comp_.size()) comp_.resize(comp_.size() * 2);  // cannot happen (block <= 64 KiB)
        if (blocks.empty() && pos > 0) {  // only empty blocks so far: drop them and refill
            memmove(comp_.data(), comp_.data() + pos, comp_len_ - pos);
            comp_len_ -= pos;
            pos = 0;
            continue;
        }
        break;  // inflate what we have; the partial block stays for the next call
    }
    if (c->head + out > c->cap) {
        err_ = SPH_EFORMAT;
        err_text_ = "internal: chunk overflow";
        return false;
    }
    // parallel inflate, ~32 blocks per task
    const int64_t per = 32;
    const int64_t nb = (int64_t) blocks.size();
    std::atomic<int> bad{0};
    const uint8_t *comp = comp_.data();
    uint8_t *dst = c->buf;
    static const bool use_zlib_only = getenv("SPH_ZLIB_INFLATE") != nullptr;
    const size_t comp_cap = comp_.size();
    pool_->parallel_for((nb + per - 1) / per, [&](int64_t t) {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) {
            bad = 1;
            return;
        }
        for (int64_t i = t * per; i < std::min(nb, (t + 1) * per); i++) {
            const BlockRef &r = blocks[(size_t) i];
            // fast path (sph_inflate.cpp), accepted only with a matching CRC32; anything else goes to zlib below
            if (!use_zlib_only && r.comp_off + r.comp_len + 8 <= comp_cap &&
                fast_inflate(comp + r.comp_off, r.comp_len, dst + r.uoff, r.ulen) &&
                (uint32_t) crc32(crc32(0L, Z_NULL, 0), dst + r.uoff, r.ulen) == r.crc)
                continue;
            zs.next_in = const_cast<Bytef *>(comp + r.comp_off);
            zs.avail_in = r.comp_len;
            zs.next_out = dst + r.uoff;
            zs.avail_out = r.ulen;
            int rc = inflate(&zs, Z_FINISH);
            if (rc != Z_STREAM_END || zs.avail_out != 0) bad = 1;
            else if ((uint32_t) crc32(crc32(0L, Z_NULL, 0), dst + r.uoff, r.ulen) != r.crc) bad = 2;
            inflateReset(&zs);
        }
        inflateEnd(&zs);
    });
    if (bad) {
        err_ = SPH_EFORMAT;
        err_text_ = bad == 2 ? "BGZF CRC mismatch" : "BGZF inflate failed (corrupt block)";
        return false;
    }
    memmove(comp_.data(), comp_.data() + pos, comp_len_ - pos);
    comp_len_ -= pos;
    c->end = c->head + out;
    c->last = file_eof_ && comp_len_ == 0;
    return out > 0 || c->last;
}

void BgzfReader::producer() {
    for (;;) {
        Chunk *c = nullptr;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || !free_.empty(); });
            if (stop_) return;
            c = free_.back();
            free_.pop_back();
        }
        bool ok = fill(c);
        std::lock_guard<std::mutex> g(mu_);
        if (!ok || err_) {
            free_.push_back(c);
            done_ = true;
            cv_.notify_all();
            return;
        }
        ready_.push_back(c);
        if (c->last) done_ = true;
        cv_.notify_all();
        if (done_) return;
    }
}

Chunk *BgzfReader::next() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !ready_.empty() || done_; });
    if (ready_.empty()) return nullptr;
    Chunk *c = ready_.front();
    ready_.erase(ready_.begin());
    return c;
}

void BgzfReader::release(Chunk *c) {
    std::lock_guard<std::mutex> g(mu_);
    free_.push_back(c);
    cv_.notify_all();
}

// ------------------------------------------------------------------------------------------
BgzfWriter::~BgzfWriter() {
    if (fp_) fclose(fp_);
}

int BgzfWriter::open(const char *path, int level, WorkerPool *pool) {
    fp_ = fopen(path, "wb");
    if (!fp_) {
        set_error("cannot create %s: %s", path, strerror(errno));
        return SPH_EIO;
    }
    path_ = path;
    level_ = level;
    pool_ = pool;
    return SPH_OK;
}

int BgzfWriter::write(const void *data, size_t len) {
    const uint8_t *p = (const uint8_t *) data;
    pend_.insert(pend_.end(), p, p + len);
    if (pend_.size() >= FLUSH_BYTES) return flush(false);
    return SPH_OK;
}

// Deflates the pending bytes as blocks of <= 0xff00 payload bytes, in parallel, and writes
// them in order.  Unless `all`, a trailing partial block stays pending.
int BgzfWriter::flush(bool all) {
    const size_t BLK = 0xff00;
    size_t len = pend_.size();
    size_t nblk = all ? (len + BLK - 1) / BLK : len / BLK;
    if (nblk == 0) return SPH_OK;
    const uint8_t *data = pend_.data();
    std::vector<std::vector<uint8_t>> outs(nblk);
    std::atomic<int> bad{0};
    const int level = level_;
    pool_->parallel_for((int64_t) nblk, [&](int64_t k) {
        size_t bi = (size_t) k;
        const uint8_t *src = data + bi * BLK;
        size_t n = std::min(BLK, len - bi * BLK);
        std::vector<uint8_t> &o = outs[bi];
        o.resize(18 + compressBound((uLong) n) + 16 + 8);
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
            bad = 1;
            return;
        }
        zs.next_in = const_cast<Bytef *>(src);
        zs.avail_in = (uInt) n;
        zs.next_out = o.data() + 18;
        zs.avail_out = (uInt) (o.size() - 18 - 8);
        if (deflate(&zs, Z_FINISH) != Z_STREAM_END) bad = 1;
        size_t clen = zs.total_out;
        deflateEnd(&zs);
        size_t total = 18 + clen + 8;
        if (total > 65536) {
            bad = 1;
            return;
        }
        static const uint8_t hdr[12] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0};
        memcpy(o.data(), hdr, 12);
        o[12] = 'B'; o[13] = 'C'; o[14] = 2; o[15] = 0;
        o[16] = (uint8_t) ((total - 1) & 0xff);
        o[17] = (uint8_t) ((total - 1) >> 8);
        uint32_t crc = (uint32_t) crc32(crc32(0L, Z_NULL, 0), src, (uInt) n);
        uint8_t *t = o.data() + 18 + clen;
        for (int i = 0; i < 4; i++) t[i] = (uint8_t) (crc >> (8 * i));
        for (int i = 0; i < 4; i++) t[4 + i] = (uint8_t) ((uint32_t) n >> (8 * i));
        o.resize(total);
    });
    if (!bad)
        for (size_t k = 0; k < nblk; k++)
            if (fwrite(outs[k].data(), 1, outs[k].size(), fp_) != outs[k].size()) bad = 2;
    if (bad) {
        set_error(bad == 2 ? "write to %s failed" : "deflate failed while writing %s", path_.c_str());
        return SPH_EIO;
    }
    size_t used = std::min(len, nblk * BLK);
    pend_.erase(pend_.begin(), pend_.begin() + (ptrdiff_t) used);
    return SPH_OK;
}

int BgzfWriter::close() {
    if (!fp_) return SPH_OK;
    int rc = flush(true);
    static const uint8_t eof_marker[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
                                           0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (rc == SPH_OK && fwrite(eof_marker, 1, 28, fp_) != 28) rc = SPH_EIO;
    if (fclose(fp_) != 0 && rc == SPH_OK) rc = SPH_EIO;
    fp_ = nullptr;
    if (rc == SPH_EIO) set_error("write to %s failed", path_.c_str());
    return rc;
}

}  // namespace sph
