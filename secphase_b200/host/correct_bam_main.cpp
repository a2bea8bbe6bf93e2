// correct_bam_main.cpp -- `correct_bam`: applies a Secphase phasing log (out.log) to a BAM.
//
// The consumer of the hot path's output format, kept as a drop-in next to `secphase`: same options
// (programs/src/correct_bam.c:218-233, defaults 246-250), same record-by-record decisions
// (correct_bam.c:354-380), BAM in, BAM out.  Pure host code: BGZF in/out through sph_bgzf (the
// reference uses htslib with an I/O thread pool, correct_bam.c:329-339).
//
//   phasing log   get_phased_read_table   correct_bam.c:32-88   "$\t<read>" opens a record, the "*"
//                 line is the old primary, the "@" line the selected secondary (secphase.c:32-57);
//                 the read is remembered with the "@" location unless both start at the same place
//   is_prim       correct_bam.c:90-105    a remembered read is primary exactly at that location,
//                 every other alignment of it becomes secondary; other reads keep their flag
//   mapq table    correct_bam.c:107-145, 168-186   read, contig, 1-based start, new MAPQ
//   filters       unmapped, excluded names, -p primary only, read / alignment length from the
//                 CIGAR (correct_bam.c:188-214), --maxMapq, --maxDiv on the "de" tag, -t drops aux
//
// Differences, on purpose: a record without a "de" tag has divergence 0 (the reference passes
// NULL to bam_aux2f there, correct_bam.c:375); the cs/MD tag is not required for the length
// filters (the reference's iterator exits without one, cigar_it.c:64-67; lengths are CIGAR sums
// either way).
#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/secphase_host.h"
#include "sph_bgzf.hpp"
#include "sph_common.hpp"

using namespace sph;

namespace {

struct Location {
    std::string contig;
    int32_t start;
    int mapq;
};

inline uint32_t le32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t) p[3] << 24); }
inline uint16_t le16(const uint8_t *p) { return (uint16_t) (p[0] | (p[1] << 8)); }

// strtok(line, "\t") semantics: runs of tabs are one separator, empty fields do not exist
std::vector<std::string> tab_tokens(const std::string &line) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < line.size()) {
        while (i < line.size() && line[i] == '\t') i++;
        size_t j = i;
        while (j < line.size() && line[j] != '\t') j++;
        if (j > i) out.push_back(line.substr(i, j - i));
        i = j;
    }
    return out;
}

bool read_line(FILE *fp, std::string &line) {
    line.clear();
    int c;
    bool any = false;
    while ((c = fgetc(fp)) != EOF) {
        any = true;
        if (c == '\n') return true;
        line += (char) c;
    }
    return any;
}

// correct_bam.c:32-88
bool load_phasing_log(const char *path, std::unordered_map<std::string, Location> &table) {
    if (!path) return true;
    FILE *fp = fopen(path, "r");
    if (!fp) return false;
    std::string line, read_name, contig_new, contig_old;
    int start_new = -1, start_old = -1;
    Location loc{"", -1, -1};
    while (read_line(fp, line)) {
        if (line.empty()) continue;
        if (line[0] == '$') {
            auto t = tab_tokens(line);
            read_name = t.size() > 1 ? t[1] : "";
            start_new = start_old = -1;
        } else if (line[0] == '@') {
            auto t = tab_tokens(line);
            if (t.size() < 4) continue;
            contig_new = t[2];
            start_new = atoi(t[3].c_str());
            loc = Location{contig_new, start_new, -1};
            if (start_old != -1 && (start_old != start_new || contig_old != contig_new)) table[read_name] = loc;
        } else if (line[0] == '*') {
            auto t = tab_tokens(line);
            if (t.size() < 4) continue;
            contig_old = t[2];
            start_old = atoi(t[3].c_str());
            if (start_new != -1 && (start_old != start_new || contig_old != contig_new)) table[read_name] = loc;
        }
    }
    fclose(fp);
    return true;
}

// correct_bam.c:107-145
bool load_mapq_table(const char *path, std::unordered_map<std::string, std::vector<Location>> &table) {
    if (!path) return true;
    FILE *fp = fopen(path, "r");
    if (!fp) return false;
    std::string line;
    while (read_line(fp, line)) {
        auto t = tab_tokens(line);
        if (t.size() < 4) continue;
        table[t[0]].push_back(Location{t[1], atoi(t[2].c_str()) - 1, atoi(t[3].c_str())});
    }
    fclose(fp);
    return true;
}

bool load_read_set(const char *path, std::unordered_set<std::string> &set) {  // correct_bam.c:148-166
    if (!path) return true;
    FILE *fp = fopen(path, "r");
    if (!fp) return false;
    std::string line;
    while (read_line(fp, line))
        if (!line.empty()) set.insert(line);
    fclose(fp);
    return true;
}

// sequential reader of the uncompressed BAM byte stream
struct Stream {
    BgzfReader *rd;
    std::vector<uint8_t> buf;
    size_t pos = 0;
    bool eof = false;
    // makes n bytes available at pos; false at end of stream
    bool need(size_t n) {
        while (buf.size() - pos < n) {
            if (eof) return false;
            Chunk *c = rd->next();
            if (!c) {
                eof = true;
                continue;
            }
            if (pos > 0) {
                buf.erase(buf.begin(), buf.begin() + (ptrdiff_t) pos);
                pos = 0;
            }
            buf.insert(buf.end(), c->buf + c->head, c->buf + c->end);
            const bool last = c->last;
            rd->release(c);
            if (last) eof = true;
        }
        return true;
    }
};

// bam_aux2f(bam_aux_get(b, "de")): numeric value of the tag, 0 if absent / not numeric
double aux_de(const uint8_t *aux, const uint8_t *end) {
    const uint8_t *p = aux;
    while (p + 3 <= end) {
        const uint8_t t0 = p[0], t1 = p[1], type = p[2];
        p += 3;
        size_t sz = 0;
        switch (type) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'd': sz = 8; break;
            case 'Z': case 'H': {
                const void *z = memchr(p, 0, (size_t) (end - p));
                if (!z) return 0.0;
                sz = (size_t) ((const uint8_t *) z - p) + 1;
                break;
            }
            case 'B': {
                if (end - p < 5) return 0.0;
                const uint8_t sub = p[0];
                const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                sz = 5 + es * (size_t) le32(p + 1);
                break;
            }
            default: return 0.0;
        }
        if (p + sz > end) return 0.0;
        if (t0 == 'd' && t1 == 'e') {
            switch (type) {
                case 'f': { uint32_t u = le32(p); float f; memcpy(&f, &u, 4); return f; }
                case 'd': { double d; memcpy(&d, p, 8); return d; }
                case 'c': return (int8_t) p[0];
                case 'C': return p[0];
                case 's': return (int16_t) le16(p);
                case 'S': return le16(p);
                case 'i': return (int32_t) le32(p);
                case 'I': return le32(p);
                default: return 0.0;
            }
        }
        p += sz;
    }
    return 0.0;
}

void usage(const char *program) {  // correct_bam.c:289-312
    fprintf(stderr,
            "\nUsage: %s  -i <INPUT_BAM> -o <OUTPUT_BAM> -p <PHASING_LOG> -m <MAPQ_TABLE>\n\tModify the input bam file:\n\t* Apply the phasing log by swapping the primary and secondary alignments whenever necessary(stdout log of ./phase_reads)\n\t* Set the MAPQs to the values given in the mapq table\n\t\tmapq table is a tab delimited text containing 4 columns:\n\t\t1. read name\n\t\t2. contig name\n\t\t3. left-most coordinate on contig (1-based)\n\t\t4. adjusted mapq\n\t* Filter secondary alignments (After applying the phasing log)\n\t* Skip outputing the optional fields (like cs and MD tags)\n\t* Filter the reads shorter than the given threshold\n\t* Filter the alignments shorter than the given threshold\n\n",
            program);
    fprintf(stderr, "Options:\n");
    fprintf(stderr, "         --inputBam,\t-i         input bam file\n");
    fprintf(stderr, "         --outputBam,\t-o         output bam file\n");
    fprintf(stderr, "         --maxMapq,\t-x         maximum mapq [default:100]\n");
    fprintf(stderr, "         --phasingLog,\t-P         the phasing log path (output of secphase) [optional]\n");
    fprintf(stderr,
            "         --mapqTable,\t-M         the adjusted mapq table (tab-delimited) path (4 columns with no header: read_name, contig_name, 1_based_contig_start , new_mapq) [optional]\n");
    fprintf(stderr, "         --exclude,\t-e         Path to a file containing the read names that have to be excluded [optional]\n");
    fprintf(stderr, "         --noTag,\t-t         output no optional fields\n");
    fprintf(stderr, "         --primaryOnly,\t-p         output only primary alignments\n");
    fprintf(stderr, "         --minReadLen,\t-m         min read length [default: 5k]\n");
    fprintf(stderr, "         --minAlignmentLen,\t-a         min alignment length [default: 5k]\n");
    fprintf(stderr, "         --maxDiv,\t-d         min gap-compressed divergence (\"de\" tag) [default: 0.12]\n");
    fprintf(stderr, "         --threads,\t-n         number of threads (for bam I/O)[default: 2]\n");
}

}  // namespace

int main(int argc, char *argv[]) {
    static struct option long_options[] = {{"inputBam", required_argument, nullptr, 'i'},
                                           {"outputBam", required_argument, nullptr, 'o'},
                                           {"phasingLog", required_argument, nullptr, 'P'},
                                           {"mapqTable", required_argument, nullptr, 'M'},
                                           {"minReadLen", required_argument, nullptr, 'm'},
                                           {"minAlignmentLen", required_argument, nullptr, 'a'},
                                           {"primaryOnly", no_argument, nullptr, 'p'},
                                           {"exclude", required_argument, nullptr, 'e'},
                                           {"threads", required_argument, nullptr, 'n'},
                                           {"noTag", no_argument, nullptr, 't'},
                                           {"maxMapq", required_argument, nullptr, 'x'},
                                           {"maxDiv", required_argument, nullptr, 'd'},
                                           {nullptr, 0, nullptr, 0}};
    const char *input_path = nullptr, *output_path = nullptr, *exclude_path = nullptr, *phasing_log_path = nullptr,
               *mapq_table_path = nullptr;
    bool primary_only = false, no_tag = false;
    int min_read_length = 5000, min_alignment_length = 5000, nthreads = 2, max_mapq = 100;
    double max_divergence = 0.12;
    const char *program = strrchr(argv[0], '/');
    program = program ? program + 1 : argv[0];
    int c;
    while (~(c = getopt_long(argc, argv, "i:o:x:e:P:M:tpm:a:n:d:h", long_options, nullptr))) {
        switch (c) {
            case 'i': input_path = optarg; break;
            case 'o': output_path = optarg; break;
            case 'x': max_mapq = atoi(optarg); break;
            case 'e': exclude_path = optarg; break;
            case 'P': phasing_log_path = optarg; break;
            case 'M': mapq_table_path = optarg; break;
            case 't': no_tag = true; break;
            case 'p': primary_only = true; break;
            case 'm': min_read_length = atoi(optarg); break;
            case 'a': min_alignment_length = atoi(optarg); break;
            case 'n': nthreads = atoi(optarg); break;
            case 'd': max_divergence = atof(optarg); break;
            default:
                if (c != 'h') fprintf(stderr, "[E::%s] undefined option %c\n", __func__, c);
                usage(program);
                return 1;
        }
    }
    if (!input_path || !output_path) {
        fprintf(stderr, "[E::%s] both -i <INPUT_BAM> and -o <OUTPUT_BAM> are required\n", __func__);
        usage(program);
        return 1;
    }
    std::unordered_map<std::string, Location> phased;
    std::unordered_map<std::string, std::vector<Location>> mapq_table;
    std::unordered_set<std::string> exclude;
    if (!load_phasing_log(phasing_log_path, phased)) {
        fprintf(stderr, "Error: cannot open %s\n", phasing_log_path);
        return 1;
    }
    if (!load_mapq_table(mapq_table_path, mapq_table)) {
        fprintf(stderr, "Error: cannot open %s\n", mapq_table_path);
        return 1;
    }
    if (!load_read_set(exclude_path, exclude)) {
        fprintf(stderr, "Error: cannot open %s\n", exclude_path);
        return 1;
    }

    WorkerPool pool(nthreads > 0 ? nthreads : 1);
    BgzfReader rd(&pool, (size_t) 16 << 20, 0);
    if (rd.open(input_path) != SPH_OK) {
        fprintf(stderr, "Error: %s\n", sph_last_error());
        return 1;
    }
    Stream in;
    in.rd = &rd;
    // header: magic, l_text, text, n_ref, (l_name, name, l_ref)* -- copied through unchanged (sam_hdr_write)
    std::vector<std::string> names;
    std::vector<uint8_t> header;
    {
        if (!in.need(12) || memcmp(in.buf.data() + in.pos, "BAM\1", 4) != 0) {
            fprintf(stderr, "Error: %s is not a BAM file%s%s\n", input_path, rd.error() ? ": " : "",
                    rd.error() ? rd.error_text().c_str() : "");
            return 1;
        }
        const uint32_t l_text = le32(in.buf.data() + in.pos + 4);
        if (!in.need(12 + (size_t) l_text)) { fprintf(stderr, "Error: truncated BAM header\n"); return 1; }
        const uint32_t n_ref = le32(in.buf.data() + in.pos + 8 + l_text);
        size_t o = 12 + (size_t) l_text;
        for (uint32_t i = 0; i < n_ref; i++) {
            if (!in.need(o + 4)) { fprintf(stderr, "Error: truncated BAM header\n"); return 1; }
            const uint32_t l_name = le32(in.buf.data() + in.pos + o);
            if (!in.need(o + 4 + l_name + 4)) { fprintf(stderr, "Error: truncated BAM header\n"); return 1; }
            {  // the name is NUL-terminated inside l_name bytes in a well-formed file; never read past them
                const char *nm = (const char *) in.buf.data() + in.pos + o + 4;
                names.emplace_back(nm, l_name ? strnlen(nm, l_name) : 0);
            }
            o += 4 + (size_t) l_name + 4;
        }
        header.assign(in.buf.begin() + (ptrdiff_t) in.pos, in.buf.begin() + (ptrdiff_t) (in.pos + o));
        in.pos += o;
    }
    BgzfWriter out;
    if (out.open(output_path, 6, &pool) != SPH_OK || out.write(header.data(), header.size()) != SPH_OK) {
        fprintf(stderr, "Error: %s\n", sph_last_error());
        return 1;
    }

    std::vector<uint8_t> rec;
    long long n_in = 0, n_out = 0;
    for (;;) {
        if (!in.need(4)) break;
        const uint32_t block_size = le32(in.buf.data() + in.pos);
        if (block_size < 32 || !in.need(4 + (size_t) block_size)) {
            fprintf(stderr, "Error: truncated or malformed BAM record after %lld records\n", n_in);
            return 1;
        }
        rec.assign(in.buf.begin() + (ptrdiff_t) (in.pos + 4), in.buf.begin() + (ptrdiff_t) (in.pos + 4 + block_size));
        in.pos += 4 + (size_t) block_size;
        n_in++;
        uint8_t *b = rec.data();
        const int32_t tid = (int32_t) le32(b), pos = (int32_t) le32(b + 4);
        const int l_qname = b[8];
        const int n_cigar = le16(b + 12);
        int flag = le16(b + 14);
        const int32_t l_seq = (int32_t) le32(b + 16);
        const size_t aux_off = 32 + (size_t) l_qname + 4 * (size_t) n_cigar + ((size_t) l_seq + 1) / 2 + (size_t) l_seq;
        if (l_seq < 0 || aux_off > rec.size()) {
            fprintf(stderr, "Error: malformed BAM record %lld\n", n_in);
            return 1;
        }
        if (flag & 0x4) continue;  // unmapped, correct_bam.c:355
        const std::string qname((const char *) b + 32, strnlen((const char *) b + 32, (size_t) l_qname));
        if (exclude.count(qname)) continue;
        const char *contig = (tid >= 0 && tid < (int32_t) names.size()) ? names[(size_t) tid].c_str() : "";
        // is_prim, correct_bam.c:90-105
        bool prim;
        auto ph = phased.find(qname);
        if (ph != phased.end()) prim = ph->second.contig == contig && ph->second.start == pos;
        else prim = (flag & 0x100) == 0;
        if (prim) {
            flag &= ~0x100;
        } else {
            if (primary_only) continue;
            flag |= 0x100;
        }
        // get_read_length / get_alignment_length, correct_bam.c:188-214
        long long read_len = 0, aln_len = 0;
        for (int k = 0; k < n_cigar; k++) {
            const uint32_t cg = le32(b + 32 + l_qname + 4 * k);
            const uint32_t op = cg & 15, len = cg >> 4;
            if (op == 0 || op == 7 || op == 8) { read_len += len; aln_len += len; }
            else if (op == 1 || op == 4 || op == 5) read_len += len;
        }
        if (read_len < min_read_length || aln_len < min_alignment_length) continue;
        // get_mapq, correct_bam.c:168-186
        int mapq = b[9];
        auto mq = mapq_table.find(qname);
        if (mq != mapq_table.end())
            for (const Location &l : mq->second)
                if (l.contig == contig && l.start == pos) {
                    mapq = (uint8_t) l.mapq;
                    break;
                }
        if (max_mapq < mapq) continue;
        if (max_divergence < aux_de(b + aux_off, b + rec.size())) continue;
        b[9] = (uint8_t) mapq;
        b[14] = (uint8_t) (flag & 0xff);
        b[15] = (uint8_t) ((flag >> 8) & 0xff);
        const uint32_t out_len = no_tag ? (uint32_t) aux_off : (uint32_t) rec.size();
        uint8_t bs[4] = {(uint8_t) out_len, (uint8_t) (out_len >> 8), (uint8_t) (out_len >> 16), (uint8_t) (out_len >> 24)};
        if (out.write(bs, 4) != SPH_OK || out.write(b, out_len) != SPH_OK) {
            fprintf(stderr, "Couldn't write %s\n", qname.c_str());
            return 1;
        }
        n_out++;
    }
    if (rd.error()) {
        fprintf(stderr, "Error: %s\n", rd.error_text().c_str());
        return 1;
    }
    if (out.close() != SPH_OK) {
        fprintf(stderr, "Error: %s\n", sph_last_error());
        return 1;
    }
    fprintf(stderr, "[correct_bam] %lld records read, %lld written, %zu reads re-phased by the log\n", n_in, n_out, phased.size());
    return 0;
}
