// sph_sam.cpp -- the -w/--writeBam output: alignment records as SAM text.
//
// The reference writes <outDir>/<prefix>.quality_modified.out.bam through
//   bam_fo = sam_open(output_bam_path, "w"); sam_hdr_write(bam_fo, sam_hdr);     secphase.c:651-652
//   sam_write1(bam_fo, sam_hdr, alignments[i]->record);                          secphase.c:185-187
// htslib derives the format from the MODE string, not from the file name: "w" without 'b'/'c' is
// uncompressed SAM text, so despite its name the file holds SAM lines (samtools reads it either
// way, it sniffs the content).  This file restates sam_format1's text form of a BAM record
// (htslib 1.17 sam.c; SAM spec 1.4-1.5): mandatory columns, then the aux fields as TAG:TYPE:VALUE
// with integer types folded to 'i' and floats printed with %g.
#include <cerrno>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/secphase_host.h"
#include "sph_common.hpp"

using namespace sph;

namespace {
inline uint32_t le32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t) p[3] << 24); }
inline uint16_t le16(const uint8_t *p) { return (uint16_t) (p[0] | (p[1] << 8)); }

void put_int(std::string &s, long long v) {
    char tmp[24];
    int k = snprintf(tmp, sizeof(tmp), "%lld", v);
    s.append(tmp, (size_t) k);
}
void put_g(std::string &s, double v) {
    char tmp[48];
    int k = snprintf(tmp, sizeof(tmp), "%g", v);
    s.append(tmp, (size_t) k);
}

// appends one record; false on malformed input
bool format_record(std::string &s, const uint8_t *rec, int64_t len, int32_t n_targets, const char *const *names,
                   const uint8_t *qual_override) {
    if (len < 32) return false;
    const int32_t tid = (int32_t) le32(rec), pos = (int32_t) le32(rec + 4);
    const int l_qname = rec[8], mapq = rec[9];
    const int n_cigar = le16(rec + 12), flag = le16(rec + 14);
    const int32_t l_seq = (int32_t) le32(rec + 16), mtid = (int32_t) le32(rec + 20), mpos = (int32_t) le32(rec + 24),
                  isize = (int32_t) le32(rec + 28);
    const uint8_t *qname = rec + 32;
    const uint8_t *cigar = qname + l_qname;
    const uint8_t *seq = cigar + 4 * (int64_t) n_cigar;
    const uint8_t *qual = seq + (l_seq + 1) / 2;
    const uint8_t *aux = qual + l_seq;
    const uint8_t *end = rec + len;
    if (l_seq < 0 || aux > end) return false;
    // QNAME (NUL-terminated inside the record; l_qname counts the NUL and any extra padding NULs)
    size_t qn = strnlen((const char *) qname, (size_t) l_qname);
    if (qn) s.append((const char *) qname, qn); else s += '*';
    s += '\t';
    put_int(s, flag);
    s += '\t';
    if (tid >= 0 && tid < n_targets) s += names[tid]; else s += '*';
    s += '\t';
    put_int(s, (long long) pos + 1);
    s += '\t';
    put_int(s, mapq);
    s += '\t';
    if (n_cigar == 0) s += '*';
    for (int k = 0; k < n_cigar; k++) {
        const uint32_t c = le32(cigar + 4 * k);
        put_int(s, c >> 4);
        s += "MIDNSHP=XB??????"[c & 15];
    }
    s += '\t';
    if (mtid < 0 || mtid >= n_targets) s += '*';
    else if (mtid == tid) s += '=';
    else s += names[mtid];
    s += '\t';
    put_int(s, (long long) mpos + 1);
    s += '\t';
    put_int(s, isize);
    s += '\t';
    if (l_seq == 0) {
        s += "*\t*";
    } else {
        const size_t at = s.size();
        s.resize(at + (size_t) l_seq);
        for (int32_t i = 0; i < l_seq; i++) s[at + (size_t) i] = "=ACMGRSVTWYHKDBN"[(seq[i >> 1] >> ((~i & 1) << 2)) & 15];
        s += '\t';
        const uint8_t *q = qual_override ? qual_override : qual;
        if (!qual_override && qual[0] == 0xff) {
            s += '*';
        } else {
            const size_t aq = s.size();
            s.resize(aq + (size_t) l_seq);
            for (int32_t i = 0; i < l_seq; i++) s[aq + (size_t) i] = (char) (q[i] + 33);
        }
    }
    // aux fields
    const uint8_t *p = aux;
    while (p + 3 <= end) {
        const char t0 = (char) p[0], t1 = (char) p[1];
        const uint8_t type = p[2];
        p += 3;
        s += '\t';
        s += t0;
        s += t1;
        s += ':';
        switch (type) {
            case 'A': if (p + 1 > end) return false; s += "A:"; s += (char) *p; p += 1; break;
            case 'c': if (p + 1 > end) return false; s += "i:"; put_int(s, (int8_t) *p); p += 1; break;
            case 'C': if (p + 1 > end) return false; s += "i:"; put_int(s, *p); p += 1; break;
            case 's': if (p + 2 > end) return false; s += "i:"; put_int(s, (int16_t) le16(p)); p += 2; break;
            case 'S': if (p + 2 > end) return false; s += "i:"; put_int(s, le16(p)); p += 2; break;
            case 'i': if (p + 4 > end) return false; s += "i:"; put_int(s, (int32_t) le32(p)); p += 4; break;
            case 'I': if (p + 4 > end) return false; s += "i:"; put_int(s, le32(p)); p += 4; break;
            case 'f': {
                if (p + 4 > end) return false;
                uint32_t u = le32(p);
                float f;
                memcpy(&f, &u, 4);
                s += "f:";
                put_g(s, f);
                p += 4;
                break;
            }
            case 'd': {
                if (p + 8 > end) return false;
                uint64_t u = (uint64_t) le32(p) | ((uint64_t) le32(p + 4) << 32);
                double d;
                memcpy(&d, &u, 8);
                s += "d:";
                put_g(s, d);
                p += 8;
                break;
            }
            case 'Z': case 'H': {
                const void *z = memchr(p, 0, (size_t) (end - p));
                if (!z) return false;
                s += (char) type;
                s += ':';
                s.append((const char *) p, (size_t) ((const uint8_t *) z - p));
                p = (const uint8_t *) z + 1;
                break;
            }
            case 'B': {
                if (p + 5 > end) return false;
                const uint8_t sub = p[0];
                const uint32_t n = le32(p + 1);
                p += 5;
                size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2
                            : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
                if (es == 0 || p + es * (size_t) n > end) return false;
                s += "B:";
                s += (char) sub;
                for (uint32_t k = 0; k < n; k++, p += es) {
                    s += ',';
                    switch (sub) {
                        case 'c': put_int(s, (int8_t) *p); break;
                        case 'C': put_int(s, *p); break;
                        case 's': put_int(s, (int16_t) le16(p)); break;
                        case 'S': put_int(s, le16(p)); break;
                        case 'i': put_int(s, (int32_t) le32(p)); break;
                        case 'I': put_int(s, le32(p)); break;
                        default: {
                            uint32_t u = le32(p);
                            float f;
                            memcpy(&f, &u, 4);
                            put_g(s, f);
                        }
                    }
                }
                break;
            }
            default: return false;
        }
    }
    s += '\n';
    return true;
}
}  // namespace

struct sph_samw {
    FILE *fp = nullptr;
    std::vector<std::string> names;
    std::vector<const char *> name_ptr;
    std::unique_ptr<WorkerPool> pool;  // formats the records of a batch in parallel (text is ~2.5 bytes per base)
    std::vector<std::string> parts;
};

extern "C" {

int64_t sph_format_sam_record(char *buf, int64_t cap, const uint8_t *rec, int64_t rec_len, int32_t n_targets,
                              const char *const *target_names, const uint8_t *qual_override) {
    if (!rec || (n_targets > 0 && !target_names)) return SPH_EINVAL;
    std::string s;
    if (!format_record(s, rec, rec_len, n_targets, target_names, qual_override)) {
        set_error("malformed BAM record (%lld bytes)", (long long) rec_len);
        return SPH_EFORMAT;
    }
    if (buf && (int64_t) s.size() <= cap) memcpy(buf, s.data(), s.size());
    return (int64_t) s.size();
}

sph_samw *sph_samw_open(const char *path, const sph_bam *hdr) { return sph_samw_open_mt(path, hdr, 1); }

sph_samw *sph_samw_open_mt(const char *path, const sph_bam *hdr, int threads) {
    if (!path || !hdr) return nullptr;
    FILE *fp = fopen(path, "w");
    if (!fp) {
        set_error("cannot create %s: %s", path, strerror(errno));
        return nullptr;
    }
    sph_samw *w = new sph_samw();
    w->fp = fp;
    if (threads > 1) w->pool.reset(new WorkerPool(threads));
    const int32_t n = sph_bam_n_targets(hdr);
    for (int32_t i = 0; i < n; i++) w->names.emplace_back(sph_bam_target_name(hdr, i));
    for (const std::string &nm : w->names) w->name_ptr.push_back(nm.c_str());
    // sam_hdr_write, text mode: the header text; htslib adds the @SQ lines from the binary target
    // list when the text has none
    int64_t tl = 0;
    const char *text = sph_bam_header_text(hdr, &tl);
    std::string h(text ? text : "", text ? (size_t) tl : 0);
    while (!h.empty() && h.back() == '\0') h.pop_back();
    if (!h.empty() && h.back() != '\n') h += '\n';
    const bool has_sq = h.rfind("@SQ\t", 0) == 0 || h.find("\n@SQ\t") != std::string::npos;
    if (!has_sq)
        for (int32_t i = 0; i < n; i++)
            h += "@SQ\tSN:" + w->names[(size_t) i] + "\tLN:" + std::to_string((long long) sph_bam_target_len(hdr, i)) + "\n";
    if (!h.empty() && fwrite(h.data(), 1, h.size(), fp) != h.size()) {
        set_error("write to %s failed", path);
        fclose(fp);
        delete w;
        return nullptr;
    }
    return w;
}

int sph_samw_write_batch(sph_samw *w, const sph_batch *b, const uint8_t *baq_qual) {
    if (!w || !b) return SPH_EINVAL;
    const sp_flat_batch *v = sph_batch_view(b);
    const int64_t *off = nullptr;
    const uint8_t *pool = sph_batch_records(b, &off);
    if (v->n_alns > 0 && (!pool || !off || off[v->n_alns] == 0)) {
        set_error("sph_samw_write_batch: the batch was read without sph_batch_keep_records");
        return SPH_EINVAL;
    }
    // records are formatted in chunks (in parallel when the writer has a pool) and written in order
    const int32_t chunk = 16;
    const int64_t n_chunks = ((int64_t) v->n_alns + chunk - 1) / chunk;
    w->parts.resize((size_t) n_chunks);
    std::vector<int32_t> bad((size_t) n_chunks, -1);
    auto work = [&](int64_t k) {
        std::string &out = w->parts[(size_t) k];
        out.clear();
        const int32_t a1 = std::min<int64_t>(v->n_alns, (k + 1) * chunk);
        for (int32_t a = (int32_t) (k * chunk); a < a1; a++)
            if (!format_record(out, pool + off[a], off[a + 1] - off[a], (int32_t) w->names.size(), w->name_ptr.data(),
                               baq_qual ? baq_qual + v->qual_off[a] : nullptr)) {
                bad[(size_t) k] = a;
                return;
            }
    };
    if (w->pool) w->pool->parallel_for(n_chunks, work);
    else for (int64_t k = 0; k < n_chunks; k++) work(k);
    for (int64_t k = 0; k < n_chunks; k++) {
        if (bad[(size_t) k] >= 0) {
            set_error("malformed BAM record (alignment %d of the batch)", bad[(size_t) k]);
            return SPH_EFORMAT;
        }
        const std::string &out = w->parts[(size_t) k];
        if (!out.empty() && fwrite(out.data(), 1, out.size(), w->fp) != out.size()) return SPH_EIO;
    }
    return SPH_OK;
}

int sph_samw_close(sph_samw *w) {
    if (!w) return SPH_EINVAL;
    int rc = fclose(w->fp) == 0 ? SPH_OK : SPH_EIO;
    delete w;
    return rc;
}

}  // extern "C"
