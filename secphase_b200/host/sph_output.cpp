// sph_output.cpp -- text outputs of the marker path: out.log records and the BED block tables.
//
// Restates, on plain vectors instead of sonLib stHash/stList of heap ptBlocks:
//   ptBlock_add_alignment                    ptBlock.c:551-571   (block [rfs,rfe], count 1)
//   ptMarker_add_marker_blocks_by_contig     ptMarker.c:851-865  (1-bp blocks at ref_pos)
//   ptBlock_sort_stHash_by_rfs               ptBlock.c:228-236
//   ptBlock_merge_blocks_v2                  ptBlock.c:274-428   (coverage segmentation)
//   ptBlock_merge_blocks                     ptBlock.c:238-272   (union)
//   ptBlock_get_total_length_by_rf / _number ptBlock.c:497-525
//   ptBlock_save_in_bed                      ptBlock.c:573-602
//   print_alignment_scores                   secphase.c:32-57
// The merge results are compared with the reference's own KATs (secphase_test.c:83-231) and
// with the reference functions themselves (oracle/_ref) in tests/test_host_output.py.
#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/secphase_host.h"
#include "sph_common.hpp"

using namespace sph;

namespace {
struct Blk {
    int32_t s, e, count;
};
}  // namespace

struct sph_blocks {
    bool with_count;
    // std::map iterates in byte-wise key order == the strcmp order ptBlock_save_in_bed sorts by
    std::map<std::string, std::vector<Blk>> per_contig;
};

// ptBlock_merge_blocks_v2 on one contig's blocks, already sorted by start.  `data` handling:
// copies of b1 keep b1's count; "extend" adds b2's count (extend_count_data); blocks without
// count data (marker blocks) carry count 0 and nothing is added.
static std::vector<Blk> merge_v2(const std::vector<Blk> &blocks, bool with_count) {
    std::vector<Blk> fin, ongoing, temp;
    if (blocks.empty()) return fin;
    for (size_t i = 0; i < blocks.size(); i++) {
        const Blk &b2 = blocks[i];
        if (ongoing.empty()) {
            ongoing.push_back(b2);
            continue;
        }
        const int32_t s2 = b2.s, e2 = b2.e;
        int32_t s1 = 0, e1 = 0;
        temp.swap(ongoing);
        ongoing.clear();
        for (const Blk &b1 : temp) {
            e1 = b1.e;
            s1 = b1.s;
            if (e1 < s2) {
                fin.push_back(b1);
            } else if (s1 <= s2) {
                if (s1 < s2) fin.push_back({s1, s2 - 1, b1.count});
                ongoing.push_back({s2, std::min(e1, e2), with_count ? b1.count + b2.count : b1.count});
                if (e2 < e1) ongoing.push_back({e2 + 1, e1, b1.count});
            } else if (e1 <= e2) {
                ongoing.push_back({s1, e1, with_count ? b1.count + b2.count : b1.count});
            } else {
                if (s1 <= e2) ongoing.push_back({s1, e2, with_count ? b1.count + b2.count : b1.count});
                ongoing.push_back({std::max(e2 + 1, s1), e1, b1.count});
            }
        }
        // the part of b2 past the last ongoing block
        if (std::max(e1 + 1, s2) <= e2) ongoing.push_back({std::max(e1 + 1, s2), e2, b2.count});
    }
    fin.insert(fin.end(), ongoing.begin(), ongoing.end());
    return fin;
}

static std::vector<Blk> merge_union(const std::vector<Blk> &blocks, bool with_count) {
    std::vector<Blk> out;
    for (size_t i = 0; i < blocks.size(); i++) {
        const Blk &b = blocks[i];
        if (i == 0 || out.back().e < b.s) {
            out.push_back(b);
        } else {
            out.back().e = std::max(out.back().e, b.e);
            if (with_count) out.back().count += b.count;
        }
    }
    return out;
}

extern "C" {

sph_blocks *sph_blocks_create(int with_count) {
    sph_blocks *t = new sph_blocks();
    t->with_count = with_count != 0;
    return t;
}
void sph_blocks_destroy(sph_blocks *t) { delete t; }

int sph_blocks_add_count(sph_blocks *t, const char *contig, int32_t rfs, int32_t rfe, int32_t count) {
    if (!t || !contig) return SPH_EINVAL;
    t->per_contig[contig].push_back({rfs, rfe, t->with_count ? count : 0});
    return SPH_OK;
}
int sph_blocks_add(sph_blocks *t, const char *contig, int32_t rfs, int32_t rfe) {
    return sph_blocks_add_count(t, contig, rfs, rfe, 1);
}

static void sort_by_start(std::vector<Blk> &v) {
    // ptBlock_cmp_rfs compares starts only; the segmentation below does not depend on the order
    // of equal starts (coverage is symmetric), so a stable sort is as good as sonLib's
    std::stable_sort(v.begin(), v.end(), [](const Blk &a, const Blk &b) { return a.s < b.s; });
}

int sph_blocks_merge_v2(sph_blocks *t) {
    if (!t) return SPH_EINVAL;
    for (auto &kv : t->per_contig) {
        sort_by_start(kv.second);
        kv.second = merge_v2(kv.second, t->with_count);
    }
    return SPH_OK;
}

int sph_blocks_merge(sph_blocks *t) {
    if (!t) return SPH_EINVAL;
    for (auto &kv : t->per_contig) {
        sort_by_start(kv.second);
        kv.second = merge_union(kv.second, t->with_count);
    }
    return SPH_OK;
}

int64_t sph_blocks_total_length(const sph_blocks *t) {
    // the reference accumulates in an int (ptBlock.c:510-524); wrap the same way
    int32_t total = 0;
    for (const auto &kv : t->per_contig)
        for (const Blk &b : kv.second) total = (int32_t) ((uint32_t) total + (uint32_t) (b.e - b.s + 1));
    return total;
}

int64_t sph_blocks_total_number(const sph_blocks *t) {
    int64_t n = 0;
    for (const auto &kv : t->per_contig) n += (int64_t) kv.second.size();
    return n;
}

int sph_blocks_save_bed(const sph_blocks *t, const char *path) {
    FILE *fp = fopen(path, "w");
    if (!fp) {
        set_error("Failed to open file %s: %s", path, strerror(errno));
        return SPH_EIO;
    }
    for (const auto &kv : t->per_contig)
        for (const Blk &b : kv.second) {
            if (b.e < b.s) continue;
            if (t->with_count) fprintf(fp, "%s\t%d\t%d\t%d\n", kv.first.c_str(), b.s, b.e + 1, b.count);
            else fprintf(fp, "%s\t%d\t%d\n", kv.first.c_str(), b.s, b.e + 1);
        }
    if (fclose(fp) != 0) {
        set_error("write to %s failed", path);
        return SPH_EIO;
    }
    return SPH_OK;
}

int64_t sph_blocks_export(const sph_blocks *t, int32_t *rows4, int64_t max_rows) {
    int64_t n = 0;
    int32_t ci = 0;
    for (const auto &kv : t->per_contig) {
        for (const Blk &b : kv.second) {
            if (rows4 && n < max_rows) {
                rows4[4 * n] = ci;
                rows4[4 * n + 1] = b.s;
                rows4[4 * n + 2] = b.e;
                rows4[4 * n + 3] = b.count;
            }
            n++;
        }
        ci++;
    }
    return n;
}

int64_t sph_format_marker_record(char *buf, int64_t cap, const char *qname, int32_t qname_len, int32_t n_alns,
                                 const int32_t *flag, const double *score, const char *const *contig,
                                 const int32_t *pos, const int32_t *rfe, int32_t best_idx) {
    std::string s = "#MARKER SCORE\n$\t";
    s.append(qname, (size_t) qname_len);
    s += "\n";
    char line[512];
    for (int32_t i = 0; i < n_alns; i++) {
        const char *tag = !(flag[i] & 0x100) ? "*" : (i == best_idx ? "@" : "!");
        int k = snprintf(line, sizeof(line), "%s\t%.2f\t", tag, score[i]);
        s.append(line, (size_t) k);
        s += contig[i];
        k = snprintf(line, sizeof(line), "\t%ld\t%d\n", (long) pos[i], rfe[i]);
        s.append(line, (size_t) k);
    }
    s += "\n";
    if ((int64_t) s.size() <= cap && buf) memcpy(buf, s.data(), s.size());
    return (int64_t) s.size();
}

}  // extern "C"
