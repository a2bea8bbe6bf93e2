// secphase_index_main.cpp -- `secphase_index`: writes <BAM>.secphase.index, the table of BGZF virtual
// offsets the reference's `secphase` needs before it will run with more than one thread
// (programs/src/secphase.c:352-382, "To use multi threading please run secphase_index ...").
//
// `bin/secphase` of this repository does not need the file (its reader inflates BGZF blocks in
// parallel from a single pass); the tool is kept so that pipelines which call it keep working and
// so that a BAM indexed here can be read by the reference.  Restates programs/src/secphase_index.c:
//   * file layout (secphase_index.c:113-118): int64 count, then `count` int64 virtual offsets
//     (compressed block address << 16 | offset inside the inflated block, htslib bgzf_tell);
//   * address[0] = position of the first alignment record (secphase_index.c:66-68);
//   * one more address after every `stepSize` counted records (86-91), taken AFTER the record was
//     read, i.e. the start of the next record;
//   * the counter (81-85) compares every record's name with the name of the FIRST record of the
//     file -- `read_name` is never updated in the loop -- so it counts the records that do not
//     belong to the first read, not the number of distinct reads.  Kept as it is: the consumer
//     only needs monotone offsets that split the file evenly (secphase.c:367-381);
//   * the last address is the end of the file (102-107).
// A block boundary belongs to the following block, as in htslib (bgzf_read moves block_address on
// when a block is used up), and the end of the file is file_size << 16.
#include <getopt.h>
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

namespace {

const char *get_timestamp() {  // common.c:12-20
    static char buf[64];
    time_t t = time(nullptr);
    struct tm tmv;
    localtime_r(&t, &tmv);
    strftime(buf, sizeof(buf), "%Y-%m-%d %H:%M:%S", &tmv);
    return buf;
}

inline uint32_t le32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t) p[3] << 24); }
inline uint16_t le16(const uint8_t *p) { return (uint16_t) (p[0] | (p[1] << 8)); }

// Sequential BGZF reader that knows where every inflated byte came from.
struct Bgzf {
    FILE *fp = nullptr;
    int64_t next_block = 0;   // file offset of the block that follows the buffered one
    int64_t block_addr = 0;   // file offset of the buffered block
    std::vector<uint8_t> blk;  // inflated payload of the buffered block
    size_t off = 0;            // read cursor inside blk
    bool eof = false;
    std::string err;

    // loads the next non-exhausted block; false at end of file or on error
    bool fill() {
        while (off == blk.size()) {
            if (eof) return false;
            uint8_t h[18];
            block_addr = next_block;
            size_t n = fread(h, 1, 18, fp);
            if (n == 0) { eof = true; blk.clear(); off = 0; return false; }
            if (n < 18 || h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) { err = "not a BGZF block"; eof = true; return false; }
            const size_t xlen = le16(h + 10);
            std::vector<uint8_t> extra(xlen);
            memcpy(extra.data(), h + 12, xlen < 6 ? xlen : 6);
            if (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, fp) != xlen - 6) { err = "truncated BGZF header"; eof = true; return false; }
            long bsize = -1;
            for (size_t o = 0; o + 4 <= xlen;) {
                const size_t slen = le16(extra.data() + o + 2);
                if (extra[o] == 'B' && extra[o + 1] == 'C' && slen == 2 && o + 6 <= xlen) bsize = (long) le16(extra.data() + o + 4) + 1;
                o += 4 + slen;
            }
            if (bsize < (long) (12 + xlen + 8)) { err = "BGZF block without a BC field"; eof = true; return false; }
            const size_t clen = (size_t) bsize - 12 - xlen;  // deflate stream + CRC32 + ISIZE
            std::vector<uint8_t> comp(clen);
            if (fread(comp.data(), 1, clen, fp) != clen) { err = "truncated BGZF block"; eof = true; return false; }
            next_block = block_addr + bsize;
            const uint32_t isize = le32(comp.data() + clen - 4), crc = le32(comp.data() + clen - 8);
            blk.resize(isize);
            off = 0;
            if (isize) {
                z_stream zs;
                memset(&zs, 0, sizeof(zs));
                if (inflateInit2(&zs, -15) != Z_OK) { err = "zlib init failed"; eof = true; return false; }
                zs.next_in = comp.data();
                zs.avail_in = (uInt) (clen - 8);
                zs.next_out = blk.data();
                zs.avail_out = isize;
                const int rc = inflate(&zs, Z_FINISH);
                inflateEnd(&zs);
                if (rc != Z_STREAM_END || zs.avail_out != 0 || crc32(0L, blk.data(), isize) != crc) {
                    err = "corrupt BGZF block";
                    eof = true;
                    return false;
                }
            }
        }
        return true;
    }
    // bgzf_tell: a used-up block hands over to the next one; end of file = file size
    int64_t tell() {
        if (off == blk.size()) return next_block << 16;
        return (block_addr << 16) | (int64_t) off;
    }
    // reads exactly n bytes; 0 = clean end of file before the first byte, -1 = truncated / error
    int read(uint8_t *dst, size_t n) {
        size_t got = 0;
        while (got < n) {
            if (!fill()) return got == 0 && err.empty() ? 0 : -1;
            const size_t k = std::min(n - got, blk.size() - off);
            memcpy(dst + got, blk.data() + off, k);
            off += k;
            got += k;
        }
        return 1;
    }
};

}  // namespace

int main(int argc, char *argv[]) {
    static struct option long_options[] = {{"inputBam", required_argument, nullptr, 'i'},
                                           {"stepSize", required_argument, nullptr, 's'},
                                           {nullptr, 0, nullptr, 0}};
    int step_size = 10;  // secphase_index.c:40
    std::string input;
    const char *program = strrchr(argv[0], '/');
    program = program ? program + 1 : argv[0];
    int c;
    while (~(c = getopt_long(argc, argv, "i:s:h", long_options, nullptr))) {
        switch (c) {
            case 'i': input = optarg; break;
            case 's': step_size = atoi(optarg); break;
            default:
                if (c != 'h') fprintf(stderr, "[E::%s] undefined option %c\n", __func__, c);
                fprintf(stderr, "\nUsage: %s  -i <INPUT_BAM> \n", program);
                fprintf(stderr, "Options:\n");
                fprintf(stderr, "         --inputBam, -i         Input BAM file\n");
                fprintf(stderr, "         --stepSize, -s         Step size for indexing\n");
                return 1;
        }
    }
    if (input.empty() || step_size < 1) {
        fprintf(stderr, "\nUsage: %s  -i <INPUT_BAM> \n", program);
        return 1;
    }
    Bgzf bz;
    bz.fp = fopen(input.c_str(), "rb");
    if (!bz.fp) {
        fprintf(stderr, "[%s] Error: cannot open %s\n", get_timestamp(), input.c_str());
        return 1;
    }
    // header (sam_hdr_read)
    std::vector<uint8_t> tmp(12);
    if (bz.read(tmp.data(), 12) != 1 || memcmp(tmp.data(), "BAM\1", 4) != 0) {
        fprintf(stderr, "[%s] Error: %s is not a BAM file\n", get_timestamp(), input.c_str());
        return 1;
    }
    {
        const uint32_t l_text = le32(tmp.data() + 4);
        // the 4 bytes read after l_text above were the first bytes of the text (or n_ref when l_text == 0)
        std::vector<uint8_t> rest;
        rest.insert(rest.end(), tmp.begin() + 8, tmp.end());
        size_t need = (size_t) l_text + 4;  // text + n_ref
        rest.resize(need > 4 ? need : 4);
        if (need > 4 && bz.read(rest.data() + 4, need - 4) != 1) { fprintf(stderr, "[%s] Error: truncated BAM header\n", get_timestamp()); return 1; }
        const uint32_t n_ref = le32(rest.data() + l_text);
        for (uint32_t i = 0; i < n_ref; i++) {
            uint8_t ln[4];
            if (bz.read(ln, 4) != 1) { fprintf(stderr, "[%s] Error: truncated BAM header\n", get_timestamp()); return 1; }
            std::vector<uint8_t> nm((size_t) le32(ln) + 4);
            if (bz.read(nm.data(), nm.size()) != 1) { fprintf(stderr, "[%s] Error: truncated BAM header\n", get_timestamp()); return 1; }
        }
    }
    std::vector<int64_t> addresses;
    addresses.push_back(bz.tell());  // the first alignment, secphase_index.c:66-68
    std::string first_name;
    bool have_first = false;
    long long count_parsed_reads = 0, step_index = 1, step_log_index = 1;
    const long long step_log_size = 100000;
    std::vector<uint8_t> rec;
    for (;;) {
        uint8_t bs[4];
        const int r = bz.read(bs, 4);
        if (r <= 0) {
            if (r < 0) { fprintf(stderr, "[%s] Error: %s\n", get_timestamp(), bz.err.empty() ? "truncated BAM record" : bz.err.c_str()); return 1; }
            break;
        }
        const uint32_t block_size = le32(bs);
        rec.resize(block_size);
        if (block_size < 32 || bz.read(rec.data(), block_size) != 1) {
            fprintf(stderr, "[%s] Error: truncated BAM record\n", get_timestamp());
            return 1;
        }
        const size_t l_qname = std::min<size_t>(rec[8], block_size - 32);  // never past the record
        const std::string name((const char *) rec.data() + 32, strnlen((const char *) rec.data() + 32, l_qname));
        if (!have_first) {
            first_name = name;
            have_first = true;
        }
        if (name != first_name) count_parsed_reads += 1;  // secphase_index.c:81-85 (see the header of this file)
        if (count_parsed_reads == step_index * step_size) {
            addresses.push_back(bz.tell());
            step_index += 1;
        }
        if (count_parsed_reads == step_log_index * step_log_size) {
            fprintf(stderr, "[%s] # Parsed reads = %lld\n", get_timestamp(), count_parsed_reads);
            step_log_index += 1;
        }
    }
    if (!bz.err.empty()) {
        fprintf(stderr, "[%s] Error: %s\n", get_timestamp(), bz.err.c_str());
        return 1;
    }
    addresses.push_back(bz.tell());  // end of the file, secphase_index.c:102-107
    fprintf(stderr, "[%s] # Total parsed reads = %lld\n", get_timestamp(), count_parsed_reads);
    const std::string index_path = input + ".secphase.index";
    fprintf(stderr, "[%s] Writing index file %s\n", get_timestamp(), index_path.c_str());
    FILE *out = fopen(index_path.c_str(), "wb");
    if (!out) {
        fprintf(stderr, "[%s] Error: cannot create %s\n", get_timestamp(), index_path.c_str());
        return 1;
    }
    const int64_t n = (int64_t) addresses.size();
    fwrite(&n, sizeof(int64_t), 1, out);
    fwrite(addresses.data(), sizeof(int64_t), addresses.size(), out);
    fclose(out);
    fclose(bz.fp);
    fprintf(stderr, "[%s] Done!\n", get_timestamp());
    return 0;
}
