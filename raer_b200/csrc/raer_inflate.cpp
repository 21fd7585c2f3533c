#include "raer_inflate.h"

#include <string.h>

namespace raer {
namespace {

enum { LL_BITS = 11, D_BITS = 8, LL_SIZE = (1 << LL_BITS) + 2048, D_SIZE = (1 << D_BITS) + 1024 };
// table entry: bits 0-4 code length to consume (sub-table entries: length beyond the primary bits),
// bits 5-7 kind, bits 8-12 number of extra bits (or sub-table index bits), bits 16-31 value
enum { K_LIT = 0, K_LEN = 1, K_EOB = 2, K_SUB = 3, K_BAD = 4 };
static inline uint32_t mk(int len, int kind, int extra, int value) {
    return (uint32_t)len | ((uint32_t)kind << 5) | ((uint32_t)extra << 8) | ((uint32_t)value << 16);
}
static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

static inline uint32_t rev_bits(uint32_t c, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (c & 1); c >>= 1; }
    return r;
}

// Canonical Huffman decode table. what: 0 literal/length alphabet, 1 distance alphabet, 2 code-length alphabet.
// Returns the number of table entries used, or -1 for an over-subscribed / unusable code.
static int build_table(const uint8_t* lens, int n, int tbits, uint32_t* tab, int cap, int what) {
    int count[16] = {0};
    for (int i = 0; i < n; ++i) ++count[lens[i]];
    count[0] = 0;
    int left = 1, used_syms = 0;
    for (int l = 1; l <= 15; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return -1;
        used_syms += count[l];
    }
    const int primary = 1 << tbits;
    for (int i = 0; i < primary; ++i) tab[i] = mk(1, K_BAD, 0, 0);
    if (used_syms == 0) return primary;
    // incomplete codes are legal only with a single code (RFC 1951 / zlib accept one distance code of one bit)
    if (left > 0 && !(used_syms == 1)) return -1;
    uint32_t next_code[16];
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + count[l - 1]) << 1; next_code[l] = code; }
    // longest code below each primary prefix -> sub-table sizes
    uint8_t sub_bits[1 << LL_BITS];
    memset(sub_bits, 0, (size_t)primary);
    uint32_t codes[320];
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = rev_bits(next_code[l]++, l);
        codes[s] = r;
        if (l > tbits) {
            const uint32_t p = r & (primary - 1);
            if (l - tbits > sub_bits[p]) sub_bits[p] = (uint8_t)(l - tbits);
        }
    }
    int used = primary;
    for (int p = 0; p < primary; ++p)
        if (sub_bits[p]) {
            const int sz = 1 << sub_bits[p];
            if (used + sz > cap) return -1;
            tab[p] = mk(tbits, K_SUB, sub_bits[p], used);
            for (int i = 0; i < sz; ++i) tab[used + i] = mk(1, K_BAD, 0, 0);
            used += sz;
        }
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        uint32_t e;
        if (what == 0) {
            if (s < 256) e = mk(0, K_LIT, 0, s);
            else if (s == 256) e = mk(0, K_EOB, 0, 0);
            else if (s <= 285) e = mk(0, K_LEN, kLenExtra[s - 257], kLenBase[s - 257]);
            else e = mk(0, K_BAD, 0, 0);
        } else if (what == 1) {
            e = s < 30 ? mk(0, K_LEN, kDistExtra[s], kDistBase[s]) : mk(0, K_BAD, 0, 0);
        } else {
            e = mk(0, K_LIT, 0, s);
        }
        const uint32_t r = codes[s];
        if (l <= tbits) {
            e |= (uint32_t)l;
            for (uint32_t i = r; i < (uint32_t)primary; i += 1u << l) tab[i] = e;
        } else {
            const uint32_t p = r & (primary - 1);
            const int sb = sub_bits[p];
            const int base = (int)(tab[p] >> 16);
            e |= (uint32_t)(l - tbits);
            for (uint32_t i = r >> tbits; i < (1u << sb); i += 1u << (l - tbits)) tab[base + i] = e;
        }
    }
    return used;
}

struct Tables {
    uint32_t ll[LL_SIZE];
    uint32_t d[D_SIZE];
};

static inline uint64_t load64(const uint8_t* p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;  // little-endian hosts only (x86-64, aarch64)
}

}  // namespace

int inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
    const uint8_t* ip = in;
    const uint8_t* const in_end = in + in_len;  // 16 readable bytes follow (BGZF footer + the next block's header)
    uint8_t* op = out;
    uint8_t* const out_end = out + out_len;
    uint64_t bb = 0;  // bit buffer, LSB first
    int bc = 0;       // valid bits in bb
    Tables T;

#define REFILL()                                                       \
    do {                                                               \
        if (ip > in_end + 8) return -1;                                \
        bb |= load64(ip) << bc;                                        \
        ip += (63 - bc) >> 3;                                          \
        bc |= 56;                                                      \
    } while (0)
#define DROP(n) do { bb >>= (n); bc -= (n); } while (0)

    int last;
    do {
        REFILL();
        last = (int)(bb & 1);
        const int type = (int)((bb >> 1) & 3);
        DROP(3);
        if (type == 0) {  // stored
            // back to the byte boundary: the bytes still in the buffer have not been consumed
            DROP(bc & 7);
            ip -= bc >> 3;
            bb = 0; bc = 0;
            if (ip + 4 > in_end) return -1;
            const unsigned len = ip[0] | (ip[1] << 8), nlen = ip[2] | (ip[3] << 8);
            ip += 4;
            if ((len ^ 0xffffu) != nlen) return -1;
            if (ip + len > in_end || op + len > out_end) return -1;
            memcpy(op, ip, len);
            ip += len;
            op += len;
            continue;
        }
        if (type == 3) return -1;
        if (type == 1) {  // fixed codes
            uint8_t lens[288 + 32];
            int i = 0;
            for (; i < 144; ++i) lens[i] = 8;
            for (; i < 256; ++i) lens[i] = 9;
            for (; i < 280; ++i) lens[i] = 7;
            for (; i < 288; ++i) lens[i] = 8;
            if (build_table(lens, 288, LL_BITS, T.ll, LL_SIZE, 0) < 0) return -1;
            for (i = 0; i < 32; ++i) lens[i] = 5;
            if (build_table(lens, 32, D_BITS, T.d, D_SIZE, 1) < 0) return -1;
        } else {  // dynamic codes
            const int hlit = (int)(bb & 31) + 257, hdist = (int)((bb >> 5) & 31) + 1, hclen = (int)((bb >> 10) & 15) + 4;
            DROP(14);
            if (hlit > 286 || hdist > 30) return -1;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t cl[19] = {0};
            REFILL();
            for (int i = 0; i < hclen; ++i) {
                if (bc < 3) REFILL();
                cl[order[i]] = (uint8_t)(bb & 7);
                DROP(3);
            }
            uint32_t cltab[1 << 7];
            if (build_table(cl, 19, 7, cltab, 1 << 7, 2) < 0) return -1;
            uint8_t lens[286 + 30 + 140];
            int n = 0;
            const int total = hlit + hdist;
            while (n < total) {
                REFILL();
                const uint32_t e = cltab[bb & 127];
                if (((e >> 5) & 7) != K_LIT) return -1;
                DROP((int)(e & 31));
                const int sym = (int)(e >> 16);
                if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
                int rep, val = 0;
                if (sym == 16) {
                    if (n == 0) return -1;
                    val = lens[n - 1];
                    rep = 3 + (int)(bb & 3); DROP(2);
                } else if (sym == 17) {
                    rep = 3 + (int)(bb & 7); DROP(3);
                } else {
                    rep = 11 + (int)(bb & 127); DROP(7);
                }
                if (n + rep > total) return -1;
                memset(lens + n, val, (size_t)rep);
                n += rep;
            }
            if (lens[256] == 0) return -1;
            if (build_table(lens, hlit, LL_BITS, T.ll, LL_SIZE, 0) < 0) return -1;
            if (build_table(lens + hlit, hdist, D_BITS, T.d, D_SIZE, 1) < 0) return -1;
        }
        // ---- the block's symbols
        while (true) {
            REFILL();  // >= 56 bits: a length (15 + 5) and a distance (15 + 13) fit without another refill
            uint32_t e = T.ll[bb & ((1u << LL_BITS) - 1)];
            int kind = (int)((e >> 5) & 7);
            if (kind == K_SUB) {
                DROP(LL_BITS);
                e = T.ll[(e >> 16) + (bb & ((1u << ((e >> 8) & 31)) - 1))];
                kind = (int)((e >> 5) & 7);
            }
            DROP((int)(e & 31));
            if (kind == K_LIT) {
                if (op >= out_end) return -1;
                *op++ = (uint8_t)(e >> 16);
                // a second and third literal from the bits already in the buffer (literal runs dominate BAM data)
                e = T.ll[bb & ((1u << LL_BITS) - 1)];
                if (((e >> 5) & 7) != K_LIT || bc < 32) continue;
                DROP((int)(e & 31));
                if (op >= out_end) return -1;
                *op++ = (uint8_t)(e >> 16);
                e = T.ll[bb & ((1u << LL_BITS) - 1)];
                if (((e >> 5) & 7) != K_LIT) continue;
                DROP((int)(e & 31));
                if (op >= out_end) return -1;
                *op++ = (uint8_t)(e >> 16);
                continue;
            }
            if (kind == K_EOB) break;
            if (kind != K_LEN) return -1;
            const int xl = (int)((e >> 8) & 31);
            const unsigned len = (e >> 16) + (unsigned)(bb & ((1u << xl) - 1));
            DROP(xl);
            uint32_t de = T.d[bb & ((1u << D_BITS) - 1)];
            if (((de >> 5) & 7) == K_SUB) {
                DROP(D_BITS);
                de = T.d[(de >> 16) + (bb & ((1u << ((de >> 8) & 31)) - 1))];
            }
            if (((de >> 5) & 7) != K_LEN) return -1;
            DROP((int)(de & 31));
            const int xd = (int)((de >> 8) & 31);
            const unsigned dist = (de >> 16) + (unsigned)(bb & ((1u << xd) - 1));
            DROP(xd);
            if (bc < 0) return -1;
            if (dist > (size_t)(op - out) || len > (size_t)(out_end - op)) return -1;
            const uint8_t* src = op - dist;
            if (dist >= 8 && (size_t)(out_end - op) >= len + 8) {
                uint8_t* const stop = op + len;
                do {
                    memcpy(op, src, 8);
                    op += 8; src += 8;
                } while (op < stop);
                op = stop;
            } else if (dist == 1) {
                memset(op, *src, len);
                op += len;
            } else {
                for (unsigned i = 0; i < len; ++i) op[i] = src[i];
                op += len;
            }
        }
    } while (!last);
#undef REFILL
#undef DROP
    // every output byte produced, and the input not overrun (the bits still buffered belong to the payload's tail)
    if (op != out_end) return -1;
    if (ip - (bc >> 3) > in_end) return -1;
    return 0;
}

}  // namespace raer
