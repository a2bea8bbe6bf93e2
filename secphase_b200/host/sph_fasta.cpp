// sph_fasta.cpp -- whole-assembly FASTA load into 1 byte/base codes.
//
// The reference opens the .fai once per JOB (fai_load, secphase.c:101) and seeks + reads the
// window of every consensus block from disk (fai_fetch, ptMarker.c:736-744), then maps it with
// seq_nt16_int[seq_nt16_table[c]] (ptMarker.c:744).  Here the assembly is decoded once, in
// parallel, and handed to sp_set_reference_codes() to live in HBM.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/secphase_host.h"
#include "sph_common.hpp"

using namespace sph;

struct sph_fasta {
    std::vector<std::string> names;
    std::vector<int64_t> off;  // [n+1]
    std::vector<uint8_t> codes;
};

namespace {
struct CodeTable {
    uint8_t t[256];
    CodeTable() {
        memset(t, 4, sizeof(t));
        t['A'] = t['a'] = 0;
        t['C'] = t['c'] = 1;
        t['G'] = t['g'] = 2;
        t['T'] = t['t'] = 3;
        // seq_nt16_table maps the digits '0'..'3' to 1,2,4,8 as well
        t['0'] = 0; t['1'] = 1; t['2'] = 2; t['3'] = 3;
    }
};
const CodeTable kCodes;

struct Rec {
    size_t name_beg, name_end;  // header line without '>' up to the first white space
    size_t seq_beg, seq_end;    // byte range of the sequence lines
};
}  // namespace

extern "C" {

sph_fasta *sph_fasta_load(const char *path, int threads) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) {
        set_error("cannot open %s: %s", path, strerror(errno));
        return nullptr;
    }
    struct stat st;
    if (fstat(fd, &st) != 0) {
        set_error("cannot stat %s", path);
        close(fd);
        return nullptr;
    }
    size_t len = (size_t) st.st_size;
    const uint8_t *d = nullptr;
    if (len) {
        d = (const uint8_t *) mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
        if (d == MAP_FAILED) {
            set_error("cannot map %s: %s", path, strerror(errno));
            close(fd);
            return nullptr;
        }
        madvise((void *) d, len, MADV_SEQUENTIAL);
    }
    close(fd);
    if (len >= 2 && d[0] == 31 && d[1] == 139) {
        set_error("%s is compressed; a plain-text FASTA is required", path);
        munmap((void *) d, len);
        return nullptr;
    }
    // pass 1: record boundaries ('>' at line start)
    std::vector<Rec> recs;
    size_t p = 0;
    while (p < len) {
        const uint8_t *nl = (const uint8_t *) memchr(d + p, '\n', len - p);
        size_t le = nl ? (size_t) (nl - d) : len;
        if (d[p] == '>') {
            if (!recs.empty()) recs.back().seq_end = p;
            Rec r;
            r.name_beg = p + 1;
            size_t e = r.name_beg;
            while (e < le && d[e] != ' ' && d[e] != '\t' && d[e] != '\r') e++;
            r.name_end = e;
            r.seq_beg = le < len ? le + 1 : len;
            r.seq_end = len;
            recs.push_back(r);
            p = r.seq_beg;
            // jump to the next header quickly: look for "\n>"
            for (;;) {
                const uint8_t *gt = (const uint8_t *) memchr(d + p, '>', len - p);
                if (!gt) {
                    p = len;
                    break;
                }
                size_t q = (size_t) (gt - d);
                if (q == 0 || d[q - 1] == '\n') {
                    p = q;
                    break;
                }
                p = q + 1;
            }
        } else if (recs.empty()) {
            if (le > p && d[p] != '\r') {
                set_error("%s: not a FASTA file (first line does not start with '>')", path);
                munmap((void *) d, len);
                return nullptr;
            }
            p = le + 1;
        } else {
            p = le + 1;
        }
    }
    WorkerPool pool(threads > 0 ? threads : 1);
    // pass 2: bases (bytes that are not line terminators) per ~16 MB slice of every record
    struct Slice { size_t rec, beg, end; int64_t n, out; };
    std::vector<Slice> slices;
    const size_t SL = (size_t) 16 << 20;
    for (size_t i = 0; i < recs.size(); i++)
        for (size_t q = recs[i].seq_beg; q < recs[i].seq_end || q == recs[i].seq_beg; q += SL) {
            slices.push_back({i, q, std::min(q + SL, recs[i].seq_end), 0, 0});
            if (recs[i].seq_end <= recs[i].seq_beg) break;
        }
    pool.parallel_for((int64_t) slices.size(), [&](int64_t si) {
        Slice &sl = slices[(size_t) si];
        int64_t n = 0;
        for (size_t q = sl.beg; q < sl.end; q++) n += (d[q] != '\n') & (d[q] != '\r');
        sl.n = n;
    });
    sph_fasta *f = new sph_fasta();
    f->off.push_back(0);
    {
        size_t si = 0;
        for (size_t i = 0; i < recs.size(); i++) {
            int64_t n = 0;
            for (; si < slices.size() && slices[si].rec == i; si++) {
                slices[si].out = f->off.back() + n;
                n += slices[si].n;
            }
            f->names.emplace_back((const char *) d + recs[i].name_beg, recs[i].name_end - recs[i].name_beg);
            f->off.push_back(f->off.back() + n);
        }
    }
    f->codes.resize((size_t) f->off.back() + 16);
    // pass 3: translate every slice to its place
    uint8_t *out = f->codes.data();
    pool.parallel_for((int64_t) slices.size(), [&](int64_t si) {
        const Slice &sl = slices[(size_t) si];
        uint8_t *o = out + sl.out;
        for (size_t q = sl.beg; q < sl.end;) {
            const uint8_t *nl = (const uint8_t *) memchr(d + q, '\n', sl.end - q);
            size_t e = nl ? (size_t) (nl - d) : sl.end;
            size_t l = e - q;
            const uint8_t *s = d + q;
            if (l && memchr(s, '\r', l)) {
                for (size_t k = 0; k < l; k++)
                    if (s[k] != '\r') *o++ = kCodes.t[s[k]];
            } else {
                for (size_t k = 0; k < l; k++) o[k] = kCodes.t[s[k]];
                o += l;
            }
            q = e + 1;
        }
    });
    if (len) munmap((void *) d, len);
    return f;
}

void sph_fasta_free(sph_fasta *f) { delete f; }
int32_t sph_fasta_n(const sph_fasta *f) { return (int32_t) f->names.size(); }
const char *sph_fasta_name(const sph_fasta *f, int32_t i) { return f->names[(size_t) i].c_str(); }
int64_t sph_fasta_len(const sph_fasta *f, int32_t i) { return f->off[(size_t) i + 1] - f->off[(size_t) i]; }
const uint8_t *sph_fasta_codes(const sph_fasta *f) { return f->codes.data(); }
const int64_t *sph_fasta_offsets(const sph_fasta *f) { return f->off.data(); }

int sph_fasta_write(const char *path, int32_t n, const char *const *names, const char *const *seqs,
                    const int64_t *lens, int line_width) {
    FILE *fp = fopen(path, "w");
    if (!fp) {
        set_error("cannot create %s: %s", path, strerror(errno));
        return SPH_EIO;
    }
    if (line_width <= 0) line_width = 60;
    std::string fai_path = std::string(path) + ".fai";
    FILE *fai = fopen(fai_path.c_str(), "w");
    int64_t off = 0;
    for (int32_t i = 0; i < n; i++) {
        off += fprintf(fp, ">%s\n", names[i]);
        if (fai)
            fprintf(fai, "%s\t%lld\t%lld\t%d\t%d\n", names[i], (long long) lens[i], (long long) off, line_width,
                    line_width + 1);
        for (int64_t p = 0; p < lens[i]; p += line_width) {
            size_t l = (size_t) std::min<int64_t>(line_width, lens[i] - p);
            fwrite(seqs[i] + p, 1, l, fp);
            fputc('\n', fp);
            off += (int64_t) l + 1;
        }
    }
    if (fai) fclose(fai);
    if (fclose(fp) != 0) {
        set_error("write to %s failed", path);
        return SPH_EIO;
    }
    return SPH_OK;
}

}  // extern "C"
