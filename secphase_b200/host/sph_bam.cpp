// sph_bam.cpp -- BAM ingest: records -> read groups -> sp_flat_batch in (page-locked) pools.
//
// Host-side counterpart of parseAlignmentsAndScatterJobs() (reference programs/src/secphase.c:
// 230-351): sequential record scan, grouping of consecutive records by query name, the
// eligibility filter of secphase.c:285-288, and -- instead of one ptAlignment_construct +
// tpool_add_work per group (secphase.c:303,338) -- packing of the fields the marker path reads
// (core.flag/tid/pos/l_qseq/n_cigar, CIGAR, SEQ, QUAL, cs:Z or MD:Z) into the SoA batch that
// sp_submit() takes.  The scan touches one cache line per record; the bulk copies of a batch
// run on the worker pool.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/secphase_host.h"
#include "sph_bgzf.hpp"

using namespace sph;

namespace {

inline uint32_t le32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t) p[3] << 24); }
inline uint16_t le16(const uint8_t *p) { return (uint16_t) (p[0] | (p[1] << 8)); }

struct AlnDesc {
    const uint8_t *rec;  // start of the record body (after block_size)
    uint32_t body_len;
    int32_t flag, tid, pos, l_qseq, n_cigar;
    const uint8_t *cigar, *seq, *qual, *aux;
    const uint8_t *tag;  // text after the 'Z'
    int32_t tag_len, tag_kind;
    int64_t rec_index;
};

struct GroupDesc {
    const uint8_t *qname;
    int32_t qname_len;
    int32_t n;
    int64_t bytes;
    AlnDesc alns[10];  // eligible groups have 2..10 alignments (secphase.c:285-286)
};

// growable buffer from a caller-supplied allocator (page-locked when sp_host_alloc is passed)
struct Pool {
    uint8_t *p = nullptr;
    size_t cap = 0, len = 0;
};

}  // namespace

struct sph_batch {
    void *(*alloc)(size_t);
    void (*release)(void *);
    Pool cigar, tag, seq, qual;
    std::vector<int32_t> grp_aln_off, flag, tid, pos, l_qseq, n_cigar, tag_kind;
    std::vector<int64_t> qname_off, cigar_off, tag_off, seq_off, qual_off, rec_index;
    std::vector<char> qname_pool;
    sp_flat_batch view;
    // -w/--writeBam: the whole BAM record bodies (everything after block_size), needed to write the
    // records back out with modified qualities (secphase.c:182-189)
    bool keep_records = false;
    std::vector<uint8_t> rec_pool;
    std::vector<int64_t> rec_off;

    // `scale` extrapolates from what is packed so far to the full batch (>= 1): page-locked
    // allocations are slow and cudaFreeHost synchronises the device, so a pool is sized once for
    // the whole batch instead of growing chunk by chunk
    int reserve(Pool &pl, size_t need, double scale = 1.0) {
        if (need <= pl.cap) return SPH_OK;
        size_t ncap = std::max((size_t) ((double) need * scale * 1.15) + 4096, (size_t) 1 << 20);
        uint8_t *np = (uint8_t *) alloc(ncap);
        if (!np) {
            set_error("cannot allocate a %zu-byte host pool", ncap);
            return SPH_ENOMEM;
        }
        if (pl.len) memcpy(np, pl.p, pl.len);
        if (pl.p) release(pl.p);
        pl.p = np;
        pl.cap = ncap;
        return SPH_OK;
    }
    void clear() {
        cigar.len = tag.len = seq.len = qual.len = 0;
        grp_aln_off.assign(1, 0);
        qname_off.assign(1, 0);
        cigar_off.assign(1, 0);
        tag_off.assign(1, 0);
        seq_off.assign(1, 0);
        qual_off.assign(1, 0);
        flag.clear(); tid.clear(); pos.clear(); l_qseq.clear(); n_cigar.clear(); tag_kind.clear();
        rec_index.clear();
        qname_pool.clear();
        rec_pool.clear();
        rec_off.assign(1, 0);
        refresh();
    }
    int64_t bytes() const { return (int64_t) (cigar.len + tag.len + seq.len + qual.len); }
    void refresh() {
        static const uint8_t dummy[16] = {0};
        view.n_groups = (int32_t) grp_aln_off.size() - 1;
        view.n_alns = (int32_t) flag.size();
        view.grp_aln_off = grp_aln_off.data();
        view.qname_off = qname_off.data();
        view.qname_pool = qname_pool.empty() ? (const char *) dummy : qname_pool.data();
        view.flag = flag.data(); view.tid = tid.data(); view.pos = pos.data(); view.l_qseq = l_qseq.data();
        view.n_cigar = n_cigar.data(); view.tag_kind = tag_kind.data();
        view.cigar_off = cigar_off.data(); view.tag_off = tag_off.data(); view.seq_off = seq_off.data();
        view.qual_off = qual_off.data();
        view.cigar_pool = cigar.p ? (const uint32_t *) cigar.p : (const uint32_t *) dummy;
        view.tag_pool = tag.p ? (const char *) tag.p : (const char *) dummy;
        view.seq_pool = seq.p ? seq.p : dummy;
        view.qual_pool = qual.p ? qual.p : dummy;
    }
};

struct sph_bam {
    WorkerPool pool;
    BgzfReader rd;
    Chunk *cur = nullptr;
    size_t pos = 0;
    bool eof = false;
    int err = 0;
    // header
    std::string text;
    std::vector<std::string> names;
    std::vector<int64_t> lens;
    // grouping state (secphase.c:240-247)
    std::string read_name;
    bool have_name = false;
    std::vector<AlnDesc> open;
    std::vector<GroupDesc> pending;
    int64_t pending_bytes = 0;
    int64_t target_bytes = 0;  // limits of the batch being filled (pool pre-sizing)
    int32_t target_groups = 0;
    bool have_held = false;
    GroupDesc held;
    std::atomic<int64_t> n_records{0}, parsed_reads{0};
    std::vector<int64_t> limits;  // usable length per tid (empty = header lengths)
    bool final_flush_done = false;
    std::atomic<int64_t> skipped_groups{0};

    // chunk / head-room sizes can be shrunk through the environment so that tests reach the
    // carry-over and regrow paths with small files
    static size_t env_bytes(const char *name, size_t dflt) {
        const char *v = getenv(name);
        long long x = v ? atoll(v) : 0;
        return x > 0 ? (size_t) x : dflt;
    }
    explicit sph_bam(int threads)
        : pool(threads),
          rd(&pool, env_bytes("SPH_CHUNK_BYTES", (size_t) 48 << 20), env_bytes("SPH_HEAD_ROOM", (size_t) 16 << 20)) {}

    // Makes n bytes available at pos.  Pending groups must have been packed by the caller if a
    // chunk switch can happen (see ensure()).  Returns 1 ok, 0 clean EOF (no bytes at all), <0 error.
    int ensure(size_t n, sph_batch *b);
    int pack_pending(sph_batch *b);
    int close_group(bool *emitted, GroupDesc *out);
};

static int fail(sph_bam *r, int code) {
    r->err = code;
    return code;
}

int sph_bam::pack_pending(sph_batch *b) {
    if (pending.empty()) return SPH_OK;
    // destination offsets (sequential), then parallel copies
    const size_t a0 = b->flag.size();
    size_t n_new = 0;
    for (const GroupDesc &g : pending) n_new += (size_t) g.n;
    struct Dst { size_t cigar, tag, seq, qual; };
    std::vector<Dst> dst(n_new);
    std::vector<const AlnDesc *> src(n_new);
    size_t k = 0;
    size_t oc = b->cigar.len, ot = b->tag.len, os = b->seq.len, oq = b->qual.len;
    for (const GroupDesc &g : pending) {
        b->qname_pool.insert(b->qname_pool.end(), (const char *) g.qname, (const char *) g.qname + g.qname_len);
        b->qname_off.push_back((int64_t) b->qname_pool.size());
        for (int i = 0; i < g.n; i++) {
            const AlnDesc &a = g.alns[i];
            src[k] = &a;
            dst[k] = {oc, ot, os, oq};
            oc += 4 * (size_t) a.n_cigar;
            ot += (size_t) a.tag_len;
            os += (size_t) ((a.l_qseq + 1) / 2);
            oq += (size_t) a.l_qseq;
            b->flag.push_back(a.flag); b->tid.push_back(a.tid); b->pos.push_back(a.pos);
            b->l_qseq.push_back(a.l_qseq); b->n_cigar.push_back(a.n_cigar); b->tag_kind.push_back(a.tag_kind);
            b->rec_index.push_back(a.rec_index);
            if (b->keep_records) {
                b->rec_pool.insert(b->rec_pool.end(), a.rec, a.rec + a.body_len);
                b->rec_off.push_back((int64_t) b->rec_pool.size());
            }
            b->cigar_off.push_back((int64_t) (oc / 4));
            b->tag_off.push_back((int64_t) ot);
            b->seq_off.push_back((int64_t) os);
            b->qual_off.push_back((int64_t) oq);
            k++;
        }
        b->grp_aln_off.push_back((int32_t) (a0 + k));
    }
    int rc;
    // expected final size of this batch relative to what it holds after this pack
    double scale = 1.0;
    {
        const double have_b = (double) (oc + ot + os + oq), have_g = (double) (b->grp_aln_off.size() - 1);
        if (have_b > 0 && have_g > 0) {
            double by_bytes = (double) target_bytes / have_b, by_groups = (double) target_groups / have_g;
            scale = std::max(1.0, std::min(by_bytes, by_groups));
        }
    }
    if ((rc = b->reserve(b->cigar, oc + 16, scale)) || (rc = b->reserve(b->tag, ot + 16, scale)) ||
        (rc = b->reserve(b->seq, os + 16, scale)) || (rc = b->reserve(b->qual, oq + 16, scale)))
        return rc;
    uint8_t *pc = b->cigar.p, *pt = b->tag.p, *ps = b->seq.p, *pq = b->qual.p;
    pool.parallel_for((int64_t) n_new, [&](int64_t i) {
        const AlnDesc &a = *src[(size_t) i];
        const Dst &d = dst[(size_t) i];
        memcpy(pc + d.cigar, a.cigar, 4 * (size_t) a.n_cigar);
        memcpy(pt + d.tag, a.tag, (size_t) a.tag_len);
        memcpy(ps + d.seq, a.seq, (size_t) ((a.l_qseq + 1) / 2));
        memcpy(pq + d.qual, a.qual, (size_t) a.l_qseq);
    });
    b->cigar.len = oc; b->tag.len = ot; b->seq.len = os; b->qual.len = oq;
    pending.clear();
    pending_bytes = 0;
    b->refresh();
    return SPH_OK;
}

int sph_bam::ensure(size_t n, sph_batch *b) {
    while (!cur || cur->end - pos < n) {
        if (cur && cur->last) return (cur->end - pos == 0) ? 0 : SPH_EFORMAT;
        if (b) {  // descriptors of finished groups point into `cur`: copy them out first
            int rc = pack_pending(b);
            if (rc) return rc;
            if (have_held) {  // cannot happen: a held group ends the call before more is read
                set_error("internal: held group across a chunk switch");
                return SPH_EINVAL;
            }
        }
        Chunk *nx = rd.next();
        if (!nx) {
            if (rd.error()) {
                set_error("%s", rd.error_text().c_str());
                return rd.error();
            }
            if (!cur) return 0;
            return (cur->end - pos == 0) ? 0 : SPH_EFORMAT;
        }
        if (cur) {
            // carry: everything from the first kept record of the open group (or the cursor)
            const uint8_t *from = open.empty() ? cur->buf + pos : std::min(open[0].rec - 4, (const uint8_t *) cur->buf + pos);
            size_t len = (size_t) (cur->buf + cur->end - from);
            if (BgzfReader::grow_head(nx, len) != SPH_OK) {
                set_error("out of memory carrying %zu bytes across BGZF chunks", len);
                return SPH_ENOMEM;
            }
            uint8_t *to = nx->buf + nx->head - len;
            memcpy(to, from, len);
            ptrdiff_t delta = to - from;
            for (AlnDesc &a : open) {
                a.rec += delta; a.cigar += delta; a.seq += delta; a.qual += delta; a.aux += delta;
            }
            size_t new_pos = (size_t) ((cur->buf + pos + delta) - nx->buf);
            nx->head -= len;
            rd.release(cur);
            cur = nx;
            pos = new_pos;
        } else {
            cur = nx;
            pos = nx->head;
        }
    }
    return 1;
}

// size in bytes of one aux value of `type` at p (end-bounded); 0 on malformed data
static size_t aux_size(uint8_t type, const uint8_t *p, const uint8_t *end) {
    switch (type) {
        case 'A': case 'c': case 'C': return 1;
        case 's': case 'S': return 2;
        case 'i': case 'I': case 'f': return 4;
        case 'd': return 8;
        case 'Z': case 'H': {
            const void *z = memchr(p, 0, (size_t) (end - p));
            return z ? (size_t) ((const uint8_t *) z - p) + 1 : 0;
        }
        case 'B': {
            if (end - p < 5) return 0;
            size_t es;
            switch (p[0]) {
                case 'c': case 'C': es = 1; break;
                case 's': case 'S': es = 2; break;
                case 'i': case 'I': case 'f': es = 4; break;
                default: return 0;
            }
            return 5 + es * (size_t) le32(p + 1);
        }
        default: return 0;
    }
}

// bam_aux_get(b,"cs") / bam_aux_get(b,"MD") of cigar_it.c:44-45; cs preferred (cigar_it.c:46-62)
static bool find_tag(AlnDesc &a) {
    const uint8_t *p = a.aux, *end = a.rec + a.body_len;
    const uint8_t *cs = nullptr, *md = nullptr;
    size_t cs_len = 0, md_len = 0;
    while (p + 3 <= end) {
        uint8_t t0 = p[0], t1 = p[1], type = p[2];
        p += 3;
        size_t sz = aux_size(type, p, end);
        if (sz == 0 || p + sz > end) break;
        if (type == 'Z') {
            if (t0 == 'c' && t1 == 's' && !cs) { cs = p; cs_len = sz - 1; }
            else if (t0 == 'M' && t1 == 'D' && !md) { md = p; md_len = sz - 1; }
        }
        p += sz;
        if (cs) break;
    }
    if (cs) { a.tag = cs; a.tag_len = (int32_t) cs_len; a.tag_kind = 0; return true; }
    if (md) { a.tag = md; a.tag_len = (int32_t) md_len; a.tag_kind = 1; return true; }
    return false;
}

// BAM stores a CIGAR of more than 65535 operations in the tag CG:B,I and leaves the placeholder <l_seq>S<ref_len>N
// in the record; htslib's bam_read1, through which the reference reads (secphase.c:268), swaps the real CIGAR back
// in (bam_tag2cigar).  Same here: the alignment's CIGAR then points at the tag's payload.
static void restore_long_cigar(AlnDesc &a) {
    if (a.n_cigar < 1 || a.tid < 0 || a.pos < 0) return;
    const uint32_t c0 = le32(a.cigar);
    if ((c0 & 15) != 4 || (int32_t) (c0 >> 4) != a.l_qseq) return;
    const uint8_t *p = a.aux, *end = a.rec + a.body_len;
    while (p + 3 <= end) {
        const uint8_t t0 = p[0], t1 = p[1], type = p[2];
        p += 3;
        const size_t sz = aux_size(type, p, end);
        if (sz == 0 || p + sz > end) return;
        if (t0 == 'C' && t1 == 'G' && type == 'B' && sz >= 5 && (p[0] == 'I' || p[0] == 'i')) {
            const uint32_t n = le32(p + 1);
            if (n > 0 && 5 + 4 * (size_t) n <= sz) {
                a.cigar = p + 5;
                a.n_cigar = (int32_t) n;
            }
            return;
        }
        p += sz;
    }
}

// The read name changed or the file ended (secphase.c:279-315).
int sph_bam::close_group(bool *emitted, GroupDesc *out) {
    *emitted = false;
    parsed_reads++;
    int n = (int) open.size();
    int supp = 0, prim = 0;
    for (const AlnDesc &a : open) {
        if (a.flag & 0x800) supp++;
        if (!(a.flag & 0x100)) prim++;
    }
    if (n > 1 && n <= 10 && supp == 0 && prim == 1) {
        bool ok = true;
        for (AlnDesc &a : open) {
            restore_long_cigar(a);
            // the marker path is undefined for these in the reference (cigar_it.c:225-291 has no
            // case for N/P; a SEQ that disagrees with the CIGAR indexes out of bounds): skip the group
            int64_t qlen = 0, rlen = 0;
            for (int i = 0; i < a.n_cigar; i++) {
                uint32_t c = le32(a.cigar + 4 * i);
                uint32_t op = c & 15, len = c >> 4;
                if (op == 3 || op == 6 || op > 8) ok = false;
                if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) qlen += len;
                if (op == 0 || op == 2 || op == 7 || op == 8) rlen += len;
            }
            if (qlen != a.l_qseq || a.n_cigar == 0 || a.tid < 0 || a.tid >= (int32_t) names.size()) ok = false;
            if (ok) {
                int64_t lim = limits.empty() ? lens[(size_t) a.tid] : limits[(size_t) a.tid];
                if (a.pos < 0 || (int64_t) a.pos + rlen > lim) ok = false;
            }
            if (ok && a.l_qseq > 0 && a.qual[0] == 0xff) ok = false;  // QUAL absent
            if (!ok) break;
            if (!find_tag(a)) {
                set_error("record %lld (%.*s): neither cs:Z nor MD:Z present -- \"At least one of the MD or CS tags "
                          "should be present!\"", (long long) a.rec_index, (int) read_name.size(), read_name.c_str());
                return SPH_ENOTAG;
            }
        }
        if (ok) {
            out->n = n;
            out->qname = open[0].rec + 32;
            out->qname_len = (int32_t) read_name.size();
            out->bytes = 0;
            for (int i = 0; i < n; i++) {
                out->alns[i] = open[(size_t) i];
                const AlnDesc &a = open[(size_t) i];
                out->bytes += 4 * (int64_t) a.n_cigar + a.tag_len + (a.l_qseq + 1) / 2 + a.l_qseq;
            }
            *emitted = true;
        } else {
            skipped_groups++;
        }
    }
    open.clear();
    return SPH_OK;
}

extern "C" {

sph_bam *sph_bam_open(const char *path, int threads) {
    if (!path) {
        set_error("null path");
        return nullptr;
    }
    sph_bam *r = new sph_bam(std::max(1, threads));
    if (r->rd.open(path) != SPH_OK) {
        delete r;
        return nullptr;
    }
    // header: magic, l_text, text, n_ref, (l_name, name, l_ref)*
    auto bad = [&](const char *what) -> sph_bam * {
        if (!*sph::last_error() || what) set_error("%s: %s", path, what ? what : "truncated BAM header");
        delete r;
        return nullptr;
    };
    int rc = r->ensure(12, nullptr);
    if (rc <= 0) return bad(rc == 0 ? "empty file" : nullptr);
    const uint8_t *p = r->cur->buf + r->pos;
    if (memcmp(p, "BAM\1", 4) != 0) return bad("not a BAM file (bad magic)");
    uint32_t l_text = le32(p + 4);
    if (r->ensure(12 + (size_t) l_text, nullptr) <= 0) return bad(nullptr);
    p = r->cur->buf + r->pos;
    r->text.assign((const char *) p + 8, l_text);
    uint32_t n_ref = le32(p + 8 + l_text);
    r->pos += 12 + (size_t) l_text;
    for (uint32_t i = 0; i < n_ref; i++) {
        if (r->ensure(4, nullptr) <= 0) return bad(nullptr);
        uint32_t l_name = le32(r->cur->buf + r->pos);
        if (r->ensure(8 + (size_t) l_name, nullptr) <= 0) return bad(nullptr);
        p = r->cur->buf + r->pos;
        r->names.emplace_back((const char *) p + 4, l_name ? strnlen((const char *) p + 4, l_name) : 0);
        r->lens.push_back((int64_t) le32(p + 4 + l_name));
        r->pos += 8 + (size_t) l_name;
    }
    return r;
}

void sph_bam_close(sph_bam *r) { delete r; }
int32_t sph_bam_n_targets(const sph_bam *r) { return (int32_t) r->names.size(); }
const char *sph_bam_target_name(const sph_bam *r, int32_t tid) {
    return tid >= 0 && tid < (int32_t) r->names.size() ? r->names[(size_t) tid].c_str() : nullptr;
}
int64_t sph_bam_target_len(const sph_bam *r, int32_t tid) {
    return tid >= 0 && tid < (int32_t) r->lens.size() ? r->lens[(size_t) tid] : -1;
}
const char *sph_bam_header_text(const sph_bam *r, int64_t *len) {
    if (len) *len = (int64_t) r->text.size();
    return r->text.c_str();
}

sph_batch *sph_batch_create(void *(*alloc)(size_t), void (*release)(void *)) {
    sph_batch *b = new sph_batch();
    b->alloc = alloc ? alloc : malloc;
    b->release = release ? release : free;
    b->clear();
    return b;
}
void sph_batch_destroy(sph_batch *b) {
    if (!b) return;
    for (Pool *pl : {&b->cigar, &b->tag, &b->seq, &b->qual})
        if (pl->p) b->release(pl->p);
    delete b;
}
const sp_flat_batch *sph_batch_view(const sph_batch *b) { return &b->view; }
const int64_t *sph_batch_record_index(const sph_batch *b) { return b->rec_index.data(); }
void sph_batch_keep_records(sph_batch *b, int on) { b->keep_records = on != 0; }
const uint8_t *sph_batch_records(const sph_batch *b, const int64_t **off) {
    if (off) *off = b->rec_off.data();
    return b->rec_pool.data();
}

int sph_bam_set_contig_limits(sph_bam *r, const int64_t *len_by_tid) {
    if (!r || !len_by_tid) return SPH_EINVAL;
    r->limits.assign(len_by_tid, len_by_tid + r->names.size());
    return SPH_OK;
}

int64_t sph_bam_skipped_groups(const sph_bam *r) { return r->skipped_groups; }

void sph_bam_counts(const sph_bam *r, int64_t *parsed_alignments, int64_t *parsed_reads) {
    // the reference counts the final, failing sam_read1 as well (secphase.c:268-269)
    if (parsed_alignments) *parsed_alignments = r->n_records + (r->final_flush_done ? 1 : 0);
    if (parsed_reads) *parsed_reads = r->parsed_reads;
}

int32_t sph_bam_next_batch(sph_bam *r, sph_batch *b, int32_t max_groups, int64_t max_bytes) {
    if (!r || !b || max_groups < 1) {
        set_error("bad arguments");
        return SPH_EINVAL;
    }
    if (r->err) return r->err;
    b->clear();
    r->target_bytes = max_bytes;
    r->target_groups = max_groups;
    int32_t n_groups = 0;
    int64_t n_bytes = 0;
    auto take = [&](const GroupDesc &g) {
        r->pending.push_back(g);
        n_groups++;
        n_bytes += g.bytes;
    };
    if (r->have_held) {
        take(r->held);
        r->have_held = false;
    }
    int rc;
    while (!r->eof && n_groups < max_groups && n_bytes < max_bytes) {
        rc = r->ensure(4, b);
        if (rc < 0) {
            if (rc == SPH_EFORMAT && !*sph::last_error()) set_error("truncated BAM record");
            return fail(r, rc);
        }
        bool emitted = false;
        GroupDesc g;
        if (rc == 0) {  // end of file: flush the last group (secphase.c:279, bytes_read <= -1)
            r->eof = true;
            if (!r->final_flush_done) {
                r->final_flush_done = true;
                if ((rc = r->close_group(&emitted, &g)) != SPH_OK) return fail(r, rc);
                if (emitted) take(g);
            }
            break;
        }
        uint32_t body = le32(r->cur->buf + r->pos);
        if (body < 32) {
            set_error("corrupt BAM record %lld (block_size %u)", (long long) r->n_records, body);
            return fail(r, SPH_EFORMAT);
        }
        rc = r->ensure(4 + (size_t) body, b);
        if (rc <= 0) {
            set_error("truncated BAM record %lld", (long long) r->n_records);
            return fail(r, SPH_EFORMAT);
        }
        const uint8_t *rec = r->cur->buf + r->pos + 4;
        uint32_t l_read_name = rec[8];
        uint32_t n_cigar = le16(rec + 12);
        uint32_t flag = le16(rec + 14);
        uint32_t l_seq = le32(rec + 16);
        size_t fixed = 32 + (size_t) l_read_name + 4 * (size_t) n_cigar + ((size_t) l_seq + 1) / 2 + (size_t) l_seq;
        if (l_read_name == 0 || fixed > body) {
            set_error("corrupt BAM record %lld (fields exceed block_size)", (long long) r->n_records);
            return fail(r, SPH_EFORMAT);
        }
        const char *qn = (const char *) rec + 32;
        size_t qn_len = strnlen(qn, l_read_name);
        if (!r->have_name) {
            r->read_name.assign(qn, qn_len);
            r->have_name = true;
        }
        if (qn_len != r->read_name.size() || memcmp(qn, r->read_name.data(), qn_len) != 0) {
            if ((rc = r->close_group(&emitted, &g)) != SPH_OK) return fail(r, rc);
            r->read_name.assign(qn, qn_len);
            if (emitted) {
                // a group that would overflow a non-empty batch waits for the next call; the
                // current record is then re-scanned (its name now equals read_name)
                if (n_groups > 0 && n_bytes + g.bytes > max_bytes) {
                    r->held = g;
                    r->have_held = true;
                    break;
                }
                take(g);
                if (n_groups >= max_groups || n_bytes >= max_bytes) break;
            }
        }
        r->n_records++;
        if (!(flag & 0x4) && r->open.size() <= 10) {  // secphase.c:336-338
            AlnDesc a;
            a.rec = rec;
            a.body_len = body;
            a.tid = (int32_t) le32(rec);
            a.pos = (int32_t) le32(rec + 4);
            a.flag = (int32_t) flag;
            a.l_qseq = (int32_t) l_seq;
            a.n_cigar = (int32_t) n_cigar;
            a.cigar = rec + 32 + l_read_name;
            a.seq = a.cigar + 4 * (size_t) n_cigar;
            a.qual = a.seq + ((size_t) l_seq + 1) / 2;
            a.aux = a.qual + l_seq;
            a.tag = nullptr;
            a.tag_len = 0;
            a.tag_kind = 0;
            a.rec_index = r->n_records.load() - 1;
            r->open.push_back(a);
        }
        r->pos += 4 + (size_t) body;
    }
    if ((rc = r->pack_pending(b)) != SPH_OK) return fail(r, rc);
    return b->view.n_groups;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// BAM writer (test / benchmark data, and the container the reference's -w output uses)
struct sph_bamw {
    WorkerPool pool;
    BgzfWriter w;
    std::vector<uint8_t> buf;
    explicit sph_bamw(int threads) : pool(threads) {}
};

static void put32(std::vector<uint8_t> &v, uint32_t x) {
    for (int i = 0; i < 4; i++) v.push_back((uint8_t) (x >> (8 * i)));
}
static void put16(std::vector<uint8_t> &v, uint32_t x) {
    v.push_back((uint8_t) x);
    v.push_back((uint8_t) (x >> 8));
}

// UCSC binning scheme of the SAM spec (reg2bin), beg 0-based, end exclusive
static int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int) (((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int) (((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int) (((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int) (((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int) (((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

extern "C" {

sph_bamw *sph_bamw_open(const char *path, int32_t n_targets, const char *const *names, const int64_t *lens,
                        int level, int threads) {
    if (!path || n_targets < 0 || (n_targets && (!names || !lens))) {
        set_error("bad arguments");
        return nullptr;
    }
    sph_bamw *w = new sph_bamw(std::max(1, threads));
    if (w->w.open(path, level, &w->pool) != SPH_OK) {
        delete w;
        return nullptr;
    }
    std::string text = "@HD\tVN:1.6\tSO:queryname\n";
    for (int32_t i = 0; i < n_targets; i++)
        text += std::string("@SQ\tSN:") + names[i] + "\tLN:" + std::to_string((long long) lens[i]) + "\n";
    std::vector<uint8_t> &h = w->buf;
    h.insert(h.end(), {'B', 'A', 'M', 1});
    put32(h, (uint32_t) text.size());
    h.insert(h.end(), text.begin(), text.end());
    put32(h, (uint32_t) n_targets);
    for (int32_t i = 0; i < n_targets; i++) {
        size_t l = strlen(names[i]) + 1;
        put32(h, (uint32_t) l);
        h.insert(h.end(), names[i], names[i] + l);
        put32(h, (uint32_t) lens[i]);
    }
    if (w->w.write(h.data(), h.size()) != SPH_OK) {
        delete w;
        return nullptr;
    }
    h.clear();
    return w;
}

int sph_bamw_add(sph_bamw *w, const sp_flat_batch *b) {
    if (!w || !b) return SPH_EINVAL;
    std::vector<uint8_t> &v = w->buf;
    for (int32_t g = 0; g < b->n_groups; g++) {
        const char *qn = b->qname_pool + b->qname_off[g];
        size_t qn_len = (size_t) (b->qname_off[g + 1] - b->qname_off[g]);
        if (qn_len + 1 > 255) {
            set_error("query name longer than 254 characters");
            return SPH_EINVAL;
        }
        for (int32_t a = b->grp_aln_off[g]; a < b->grp_aln_off[g + 1]; a++) {
            v.clear();
            int32_t l_seq = b->l_qseq[a], n_cigar = b->n_cigar[a];
            const uint32_t *cig = b->cigar_pool + b->cigar_off[a];  // (re-pointed at the placeholder for a long CIGAR)
            int64_t rlen = 0;
            for (int i = 0; i < n_cigar; i++) {
                uint32_t op = cig[i] & 15;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += cig[i] >> 4;
            }
            size_t tag_len = (size_t) (b->tag_off[a + 1] - b->tag_off[a]);
            int kind = b->tag_kind ? b->tag_kind[a] : 0;
            // SAM spec 4.2.2: a CIGAR of more than 65535 operations goes into CG:B,I and the record carries the
            // placeholder <l_seq>S<ref_len>N (SPH_BAM_WRITE_LONG_CIGAR=<n> lowers the limit: tests of the reader)
            static const int cg_limit = getenv("SPH_BAM_WRITE_LONG_CIGAR") ? atoi(getenv("SPH_BAM_WRITE_LONG_CIGAR")) : 65535;
            const bool use_cg = n_cigar > cg_limit;
            const int n_cigar_real = n_cigar;
            uint32_t fake[2] = {((uint32_t) l_seq << 4) | 4u, ((uint32_t) rlen << 4) | 3u};
            if (use_cg) {
                cig = fake;
                n_cigar = 2;
            }
            size_t body = 32 + qn_len + 1 + 4 * (size_t) n_cigar + ((size_t) l_seq + 1) / 2 + (size_t) l_seq + 3 +
                          tag_len + 1 + (use_cg ? 8 + 4 * (size_t) n_cigar_real : 0);
            put32(v, (uint32_t) body);
            put32(v, (uint32_t) b->tid[a]);
            put32(v, (uint32_t) b->pos[a]);
            v.push_back((uint8_t) (qn_len + 1));
            v.push_back((b->flag[a] & 0x100) ? 0 : 60);  // MAPQ
            put16(v, (uint32_t) reg2bin(b->pos[a], b->pos[a] + std::max<int64_t>(rlen, 1)));
            put16(v, (uint32_t) n_cigar);
            put16(v, (uint32_t) b->flag[a]);
            put32(v, (uint32_t) l_seq);
            put32(v, (uint32_t) -1);  // next_refID
            put32(v, (uint32_t) -1);  // next_pos
            put32(v, 0);              // tlen
            v.insert(v.end(), qn, qn + qn_len);
            v.push_back(0);
            const uint8_t *cb = (const uint8_t *) cig;
            v.insert(v.end(), cb, cb + 4 * (size_t) n_cigar);
            const uint8_t *sq = b->seq_pool + b->seq_off[a];
            v.insert(v.end(), sq, sq + ((size_t) l_seq + 1) / 2);
            const uint8_t *ql = b->qual_pool + b->qual_off[a];
            v.insert(v.end(), ql, ql + l_seq);
            // tag_kind 0 = cs, 1 = MD; anything else writes an unrelated tag (tests of the
            // "neither cs nor MD" error path)
            v.push_back(kind == 0 ? 'c' : kind == 1 ? 'M' : 'X');
            v.push_back(kind == 0 ? 's' : kind == 1 ? 'D' : 'X');
            v.push_back('Z');
            const char *tg = b->tag_pool + b->tag_off[a];
            v.insert(v.end(), tg, tg + tag_len);
            v.push_back(0);
            if (use_cg) {
                v.push_back('C'); v.push_back('G'); v.push_back('B'); v.push_back('I');
                put32(v, (uint32_t) n_cigar_real);
                const uint8_t *rb = (const uint8_t *) (b->cigar_pool + b->cigar_off[a]);
                v.insert(v.end(), rb, rb + 4 * (size_t) n_cigar_real);
            }
            int rc = w->w.write(v.data(), v.size());
            if (rc != SPH_OK) return rc;
        }
    }
    return SPH_OK;
}

int sph_bamw_close(sph_bamw *w) {
    if (!w) return SPH_EINVAL;
    int rc = w->w.close();
    delete w;
    return rc;
}

int sph_bam_write(const char *path, int32_t n_targets, const char *const *names, const int64_t *lens,
                  const sp_flat_batch *b, int level, int threads) {
    sph_bamw *w = sph_bamw_open(path, n_targets, names, lens, level, threads);
    if (!w) return SPH_EIO;
    int rc = sph_bamw_add(w, b);
    int rc2 = sph_bamw_close(w);
    return rc ? rc : rc2;
}

}  // extern "C"
