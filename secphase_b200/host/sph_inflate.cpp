// sph_inflate.cpp -- raw DEFLATE (RFC 1951) decoder for whole BGZF blocks.
//
// A BGZF block is a complete deflate stream of at most 64 KiB whose inflated size (ISIZE) and CRC32 are
// known up front, so the decoder can be written for exactly that case: input and output fully in memory,
// 64-bit bit buffer refilled eight bytes at a time, 11-bit primary Huffman table with sub-tables, match
// copies in 8-byte words.  On the BAM files of this benchmark (deflate level 1 over high-entropy
// qualities) zlib's inflate runs at ~100 MB/s of output per thread and is the bound of the whole
// file-to-file run; this decoder is what replaces it on the reader's hot path (sph_bgzf.cpp).
//
// Safety net: the function never writes outside [out, out+out_len) nor reads outside [in, in+in_len+8)
// (callers keep 8 readable bytes after every block), returns false on anything it does not like, and
// the caller checks the CRC32 of what it produced; on false or a CRC mismatch the block is decoded
// again with zlib, whose verdict is the one reported.  SPH_ZLIB_INFLATE=1 in the environment bypasses it.
#include <cstdint>
#include <cstring>

#include "sph_bgzf.hpp"

namespace sph {

namespace {

constexpr int LL_BITS = 11, D_BITS = 8;
// entry: val << 16 | ebits << 8 | kind << 4 | len     (kind 0 literal, 1 base+extra, 2 end of block, 3 sub-table)
constexpr uint32_t K_LIT = 0, K_BASE = 1, K_EOB = 2, K_SUB = 3;
constexpr uint32_t F_LIT2 = 1u << 6;  // literal entry that carries two literals: val = first | second << 8, len = both codes
inline uint32_t mk(uint32_t val, uint32_t ebits, uint32_t kind, uint32_t len) { return val << 16 | ebits << 8 | kind << 4 | len; }

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

struct Tables {
    uint32_t ll[(1 << LL_BITS) + 288 * 16];
    uint32_t d[(1 << D_BITS) + 32 * 128];
};

inline uint32_t rev_bits(uint32_t c, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; i++) r |= ((c >> i) & 1u) << (n - 1 - i);
    return r;
}

// Canonical Huffman decode table from code lengths.  sym_entry(sym, len) gives the entry payload.
// Returns false for an over-subscribed or (except for the one-code case deflate allows) incomplete code.
template <class F>
bool build(const uint8_t *lens, int n_sym, int tbits, uint32_t *tab, int tab_cap, F sym_entry, bool allow_incomplete) {
    int count[16] = {0};
    for (int s = 0; s < n_sym; s++) count[lens[s]]++;
    count[0] = 0;
    int left = 1, used = 0;
    for (int l = 1; l <= 15; l++) {
        left = left * 2 - count[l];
        if (left < 0) return false;  // over-subscribed
        used += count[l];
    }
    if (left > 0 && !(allow_incomplete && used <= 1)) return false;
    uint32_t next_code[16];
    {
        uint32_t code = 0;
        for (int l = 1; l <= 15; l++) {
            code = (code + (uint32_t) count[l - 1]) << 1;
            next_code[l] = code;
        }
    }
    const int tsize = 1 << tbits;
    for (int i = 0; i < tsize; i++) tab[i] = mk(0, 0, K_EOB, 0);  // len 0 = invalid code (checked by the decoder)
    // pass 1: longest code behind every primary-table prefix that needs a sub-table
    uint8_t sub_len[1 << LL_BITS];
    memset(sub_len, 0, (size_t) tsize);
    uint32_t codes[288];
    for (int s = 0; s < n_sym; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = rev_bits(next_code[l]++, l);
        codes[s] = r;
        if (l > tbits) {
            uint8_t &m = sub_len[r & (uint32_t) (tsize - 1)];
            if (l > m) m = (uint8_t) l;
        }
    }
    int next_sub = tsize;
    for (int i = 0; i < tsize; i++)
        if (sub_len[i]) {
            const int sbits = sub_len[i] - tbits;
            if (next_sub + (1 << sbits) > tab_cap) return false;
            tab[i] = mk((uint32_t) next_sub, (uint32_t) sbits, K_SUB, (uint32_t) tbits);
            for (int k = 0; k < (1 << sbits); k++) tab[next_sub + k] = mk(0, 0, K_EOB, 0);
            next_sub += 1 << sbits;
        }
    // pass 2: fill
    for (int s = 0; s < n_sym; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = codes[s];
        if (l <= tbits) {
            const uint32_t e = sym_entry(s, (uint32_t) l);
            for (uint32_t i = r; i < (uint32_t) tsize; i += 1u << l) tab[i] = e;
        } else {
            const uint32_t pe = tab[r & (uint32_t) (tsize - 1)];
            const uint32_t base = pe >> 16, sbits = (pe >> 8) & 0xff;
            const uint32_t e = sym_entry(s, (uint32_t) (l - tbits));
            for (uint32_t i = r >> tbits; i < (1u << sbits); i += 1u << (l - tbits)) tab[base + i] = e;
        }
    }
    return true;
}

// Where two literal codes fit into the primary index together, one look-up yields both (qualities and other
// small alphabets get 4..6-bit codes, so most of a BAM block decodes two bytes per table access).
void pair_literals(uint32_t *tab) {
    static thread_local uint32_t single[1 << LL_BITS];
    memcpy(single, tab, sizeof(single));
    for (uint32_t i = 0; i < (1u << LL_BITS); i++) {
        const uint32_t e = single[i];
        const uint32_t l1 = e & 15;
        if (((e >> 4) & 3) != K_LIT || l1 == 0 || l1 >= LL_BITS) continue;
        const uint32_t e2 = single[i >> l1];  // the bits after the first code, upper index bits zero
        const uint32_t l2 = e2 & 15;
        if (((e2 >> 4) & 3) != K_LIT || l2 == 0 || l1 + l2 > LL_BITS) continue;
        tab[i] = mk((e >> 16) | ((e2 >> 16) << 8), 0, K_LIT, l1 + l2) | F_LIT2;
    }
}

inline uint32_t ll_entry(int s, uint32_t len) {
    if (s < 256) return mk((uint32_t) s, 0, K_LIT, len);
    if (s == 256) return mk(0, 0, K_EOB, len);
    if (s > 285) return mk(0, 0, K_EOB, 0);  // 286, 287 never appear in valid data
    return mk(LEN_BASE[s - 257], LEN_EXTRA[s - 257], K_BASE, len);
}
inline uint32_t d_entry(int s, uint32_t len) {
    if (s > 29) return mk(0, 0, K_EOB, 0);
    return mk(DIST_BASE[s], DIST_EXTRA[s], K_BASE, len);
}

inline uint64_t load64(const uint8_t *p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;  // little-endian hosts only (x86-64, aarch64)
}

}  // namespace

// `in` must have 8 readable bytes after in_len (they are loaded, never interpreted).
bool fast_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
    const uint8_t *ip = in, *const in_end = in + in_len;
    uint8_t *op = out, *const out_end = out + out_len;
    uint64_t bb = 0;
    int bc = 0;
    static thread_local Tables T;
    static thread_local bool fixed_ready = false;
    static thread_local Tables F;

// after REFILL at least 56 bits are valid, or the input is used up (then the stream must end within the valid bits)
#define REFILL()                                                   \
    do {                                                           \
        if (ip + 8 <= in_end + 8 && ip <= in_end) {                \
            bb |= load64(ip) << bc;                                \
            const int adv = (63 - bc) >> 3;                        \
            ip += adv;                                             \
            bc |= 56;                                              \
        }                                                          \
    } while (0)
// bits that lie beyond the end of the input were consumed: malformed
#define OVERRUN() (bc < 0 || (ip > in_end && (int64_t) (ip - in_end) * 8 > (int64_t) bc))

    bool last = false;
    while (!last) {
        REFILL();
        last = bb & 1;
        const uint32_t type = (uint32_t) (bb >> 1) & 3;
        bb >>= 3;
        bc -= 3;
        if (type == 0) {  // stored: back to a byte boundary, LEN, NLEN, bytes
            const int drop = bc & 7;
            bb >>= drop;
            bc -= drop;
            ip -= bc >> 3;  // un-read the whole bytes still in the buffer
            bb = 0;
            bc = 0;
            if (ip > in_end || in_end - ip < 4) return false;
            const uint32_t len = ip[0] | (ip[1] << 8), nlen = ip[2] | (ip[3] << 8);
            ip += 4;
            if ((len ^ nlen) != 0xffff || (size_t) (in_end - ip) < len || (size_t) (out_end - op) < len) return false;
            memcpy(op, ip, len);
            op += len;
            ip += len;
            continue;
        }
        const Tables *tb;
        if (type == 1) {
            if (!fixed_ready) {
                uint8_t l[288];
                for (int i = 0; i < 144; i++) l[i] = 8;
                for (int i = 144; i < 256; i++) l[i] = 9;
                for (int i = 256; i < 280; i++) l[i] = 7;
                for (int i = 280; i < 288; i++) l[i] = 8;
                uint8_t dl[32];
                for (int i = 0; i < 32; i++) dl[i] = 5;
                if (!build(l, 288, LL_BITS, F.ll, (int) (sizeof(F.ll) / 4), ll_entry, false)) return false;
                if (!build(dl, 32, D_BITS, F.d, (int) (sizeof(F.d) / 4), d_entry, false)) return false;
                pair_literals(F.ll);
                fixed_ready = true;
            }
            tb = &F;
        } else if (type == 2) {
            const int hlit = (int) (bb & 31) + 257, hdist = (int) ((bb >> 5) & 31) + 1, hclen = (int) ((bb >> 10) & 15) + 4;
            bb >>= 14;
            bc -= 14;
            if (hlit > 286 || hdist > 30) return false;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t cl[19] = {0};
            REFILL();
            for (int i = 0; i < hclen; i++) {
                if (bc < 3) REFILL();
                cl[order[i]] = (uint8_t) (bb & 7);
                bb >>= 3;
                bc -= 3;
            }
            uint32_t pre[1 << 7];
            if (!build(cl, 19, 7, pre, 1 << 7, [](int s, uint32_t len) { return mk((uint32_t) s, 0, K_LIT, len); }, false)) {
                // a code-length code with a single symbol is incomplete but occurs; let zlib judge
                return false;
            }
            uint8_t lens[288 + 32];
            int n = 0;
            while (n < hlit + hdist) {
                if (bc < 32) REFILL();
                const uint32_t e = pre[bb & 127];
                const uint32_t l = e & 15;
                if (l == 0) return false;
                bb >>= l;
                bc -= (int) l;
                const uint32_t s = e >> 16;
                if (s < 16) {
                    lens[n++] = (uint8_t) s;
                } else {
                    int rep;
                    uint8_t v = 0;
                    if (s == 16) {
                        if (n == 0) return false;
                        v = lens[n - 1];
                        rep = 3 + (int) (bb & 3);
                        bb >>= 2;
                        bc -= 2;
                    } else if (s == 17) {
                        rep = 3 + (int) (bb & 7);
                        bb >>= 3;
                        bc -= 3;
                    } else {
                        rep = 11 + (int) (bb & 127);
                        bb >>= 7;
                        bc -= 7;
                    }
                    if (n + rep > hlit + hdist) return false;
                    while (rep--) lens[n++] = v;
                }
                if (bc < 0) return false;
            }
            if (OVERRUN()) return false;
            if (lens[256] == 0) return false;  // no end-of-block code
            uint8_t ll_l[288], d_l[32];
            memcpy(ll_l, lens, (size_t) hlit);
            memset(ll_l + hlit, 0, (size_t) (288 - hlit));
            memcpy(d_l, lens + hlit, (size_t) hdist);
            memset(d_l + hdist, 0, (size_t) (32 - hdist));
            if (!build(ll_l, 288, LL_BITS, T.ll, (int) (sizeof(T.ll) / 4), ll_entry, false)) return false;
            if (!build(d_l, 32, D_BITS, T.d, (int) (sizeof(T.d) / 4), d_entry, true)) return false;
            pair_literals(T.ll);
            tb = &T;
        } else {
            return false;
        }
        // ---- symbols
        for (;;) {
            if (bc < 48) REFILL();  // a whole length/distance pair needs at most 15+5+15+13 bits
            uint32_t e = tb->ll[bb & ((1u << LL_BITS) - 1)];
            if (((e >> 4) & 3) == K_SUB) {
                e = tb->ll[(e >> 16) + ((bb >> LL_BITS) & ((1u << ((e >> 8) & 0xff)) - 1))];
                bb >>= LL_BITS;
                bc -= LL_BITS;
            }
            const uint32_t l = e & 15;
            if (l == 0) return false;
            bb >>= l;
            bc -= (int) l;
            const uint32_t kind = (e >> 4) & 3;
            if (kind == K_LIT) {
                // up to three look-ups without a refill (each at most 15 bits, 48 were there)
                for (int more = 2;; more--) {
                    if (e & F_LIT2) {
                        if (out_end - op < 2) return false;
                        op[0] = (uint8_t) (e >> 16);
                        op[1] = (uint8_t) (e >> 24);
                        op += 2;
                    } else {
                        if (op >= out_end) return false;
                        *op++ = (uint8_t) (e >> 16);
                    }
                    if (!more) break;
                    e = tb->ll[bb & ((1u << LL_BITS) - 1)];
                    if (((e >> 4) & 3) != K_LIT || !(e & 15)) break;
                    bb >>= e & 15;
                    bc -= (int) (e & 15);
                }
                continue;
            }
            if (kind == K_EOB) break;
            // length
            const uint32_t eb = (e >> 8) & 0xff;
            uint32_t len = (e >> 16) + (uint32_t) (bb & ((1u << eb) - 1));
            bb >>= eb;
            bc -= (int) eb;
            // distance
            uint32_t de = tb->d[bb & ((1u << D_BITS) - 1)];
            if (((de >> 4) & 3) == K_SUB) {
                de = tb->d[(de >> 16) + ((bb >> D_BITS) & ((1u << ((de >> 8) & 0xff)) - 1))];
                bb >>= D_BITS;
                bc -= D_BITS;
            }
            const uint32_t dl = de & 15;
            if (dl == 0 || ((de >> 4) & 3) != K_BASE) return false;
            bb >>= dl;
            bc -= (int) dl;
            const uint32_t deb = (de >> 8) & 0xff;
            const uint32_t dist = (de >> 16) + (uint32_t) (bb & ((1u << deb) - 1));
            bb >>= deb;
            bc -= (int) deb;
            if (bc < 0) return false;
            if (dist > (size_t) (op - out) || len > (size_t) (out_end - op)) return false;
            const uint8_t *src = op - dist;
            if (dist >= 8 && (size_t) (out_end - op) >= len + 8) {
                uint8_t *const end = op + len;
                do {
                    memcpy(op, src, 8);
                    op += 8;
                    src += 8;
                } while (op < end);
                op = end;
            } else {
                while (len--) *op++ = *src++;
            }
        }
        if (OVERRUN()) return false;
    }
#undef REFILL
#undef OVERRUN
    return op == out_end;
}

}  // namespace sph
