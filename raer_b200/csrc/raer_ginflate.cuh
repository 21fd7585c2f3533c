// raer_ginflate.cuh -- raw DEFLATE (RFC 1951) on the device, for BGZF blocks: one warp per block.
//
// A BGZF block is an independent DEFLATE stream of at most 64 KiB of output whose compressed payload and exact output
// size are known before decoding starts (BSIZE / ISIZE), so thousands of blocks decode side by side. Huffman decoding
// itself is a serial chain of table lookups: lane 0 of the warp walks the symbols, the other lanes help where the work
// is wide (clearing and filling the decode tables). The tables of the block live in the warp's slice of shared
// memory (10-bit primary table for literals / lengths, 8-bit for distances, second-level tables behind them); the
// compressed bytes are read through a 64-bit bit buffer refilled 32 bits at a time from aligned words; the output goes
// straight to the block's window in global memory, where the matches read it back from (L1 / L2 hits: the window was
// just written by the same thread). Same table layout and the same acceptance rules as the host decoder
// (raer_inflate.cpp); a block the decoder refuses is reported through `status` and the host decides what to do.
#ifndef RAER_GINFLATE_CUH
#define RAER_GINFLATE_CUH

#define GI_LL_BITS 10
#define GI_D_BITS 8
#define GI_LL_SIZE ((1 << GI_LL_BITS) + 1536)
#define GI_D_SIZE ((1 << GI_D_BITS) + 512)
#define GI_WARP_WORDS (GI_LL_SIZE + GI_D_SIZE)  // shared-memory words per warp
#define GI_WARPS 16                             // warps (= blocks in flight) per CTA, one CTA per SM

// table entry: bits 0-4 code length to consume (sub-table entries: length beyond the primary bits),
// bits 5-7 kind, bits 8-12 number of extra bits (or sub-table index bits), bits 16-31 value
enum { GK_LIT = 0, GK_LEN = 1, GK_EOB = 2, GK_SUB = 3, GK_BAD = 4 };
__device__ __forceinline__ uint32_t gi_mk(int len, int kind, int extra, int value) {
    return (uint32_t)len | ((uint32_t)kind << 5) | ((uint32_t)extra << 8) | ((uint32_t)value << 16);
}
__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// Canonical Huffman decode table, built by one thread. what: 0 literal/length alphabet, 1 distance alphabet,
// 2 code-length alphabet. Returns the number of table entries used, or -1 for an over-subscribed / unusable code or
// a table that does not fit `cap`.
__device__ int gi_build_table(const uint8_t* lens, int n, int tbits, uint32_t* tab, int cap, int what, uint8_t* sub_bits,
                              uint16_t* codes) {
    int count[16];
    for (int i = 0; i < 16; ++i) count[i] = 0;
    for (int i = 0; i < n; ++i) ++count[lens[i]];
    count[0] = 0;
    int left = 1, used_syms = 0;
    for (int l = 1; l <= 15; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return -1;
        used_syms += count[l];
    }
    const int primary = 1 << tbits;
    for (int i = 0; i < primary; ++i) tab[i] = gi_mk(1, GK_BAD, 0, 0);
    if (used_syms == 0) return primary;
    // incomplete codes are legal only with a single code (RFC 1951 / zlib accept one distance code of one bit)
    if (left > 0 && !(used_syms == 1)) return -1;
    uint32_t next_code[16];
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + count[l - 1]) << 1; next_code[l] = code; }
    for (int i = 0; i < primary; ++i) sub_bits[i] = 0;
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = __brev(next_code[l]++) >> (32 - l);
        codes[s] = (uint16_t)r;
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
            tab[p] = gi_mk(tbits, GK_SUB, sub_bits[p], used);
            for (int i = 0; i < sz; ++i) tab[used + i] = gi_mk(1, GK_BAD, 0, 0);
            used += sz;
        }
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        uint32_t e;
        if (what == 0) {
            if (s < 256) e = gi_mk(0, GK_LIT, 0, s);
            else if (s == 256) e = gi_mk(0, GK_EOB, 0, 0);
            else if (s <= 285) e = gi_mk(0, GK_LEN, c_len_extra[s - 257], c_len_base[s - 257]);
            else e = gi_mk(0, GK_BAD, 0, 0);
        } else if (what == 1) {
            e = s < 30 ? gi_mk(0, GK_LEN, c_dist_extra[s], c_dist_base[s]) : gi_mk(0, GK_BAD, 0, 0);
        } else {
            e = gi_mk(0, GK_LIT, 0, s);
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

struct GiBits {  // LSB-first bit reader over aligned 32-bit words of global memory
    const uint32_t* w;      // next word to load
    const uint32_t* w_end;  // first word that must not be loaded (a few words of slack lie behind the payload)
    unsigned long long bb;
    int bc;
    bool over;
    __device__ __forceinline__ void init(const uint8_t* p, size_t n) {
        const size_t a = (size_t)p & 3;
        w = reinterpret_cast<const uint32_t*>(p - a);
        w_end = reinterpret_cast<const uint32_t*>((((size_t)(p + n)) + 3) & ~(size_t)3) + 2;
        bb = 0;
        bc = 0;
        over = false;
        need(32);
        drop((int)a * 8);
    }
    __device__ __forceinline__ void need(int n) {  // n <= 32
        while (bc < n) {
            uint32_t v = 0;
            if (w < w_end) v = __ldg(w);
            else over = true;
            ++w;
            bb |= (unsigned long long)v << bc;
            bc += 32;
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(int n) { bb >>= n; bc -= n; }
    // bytes of the payload consumed so far (whole bytes still in the buffer are not)
    __device__ __forceinline__ const uint8_t* byte_pos() const { return reinterpret_cast<const uint8_t*>(w) - (bc >> 3); }
};

// Decodes one raw DEFLATE stream (executed by ONE thread). in / in_len: the payload; out / out_len: exactly the bytes it
// must produce. tab: GI_WARP_WORDS words of shared memory. Returns 0, or a negative code.
__device__ int gi_inflate_block(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, uint32_t* tab) {
    uint32_t* const LL = tab;
    uint32_t* const DD = tab + GI_LL_SIZE;
    GiBits B;
    B.init(in, in_len);
    size_t op = 0;
    uint8_t lens[286 + 30 + 140];
    uint8_t sub_bits[1 << GI_LL_BITS];
    uint16_t codes[320];
    int last;
    do {
        B.need(3);
        last = (int)B.peek(1);
        const int type = (int)(B.peek(3) >> 1);
        B.drop(3);
        if (type == 0) {  // stored
            B.drop(B.bc & 7);
            const uint8_t* ip = B.byte_pos();
            if (ip + 4 > in + in_len) return -2;
            const unsigned len = ip[0] | (ip[1] << 8), nlen = ip[2] | (ip[3] << 8);
            ip += 4;
            if ((len ^ 0xffffu) != nlen) return -3;
            if (ip + len > in + in_len || op + len > out_len) return -4;
            for (unsigned i = 0; i < len; ++i) out[op + i] = ip[i];
            op += len;
            B.init(ip + len, (size_t)((in + in_len) - (ip + len)));
            continue;
        }
        if (type == 3) return -5;
        if (type == 1) {  // fixed codes
            int i = 0;
            for (; i < 144; ++i) lens[i] = 8;
            for (; i < 256; ++i) lens[i] = 9;
            for (; i < 280; ++i) lens[i] = 7;
            for (; i < 288; ++i) lens[i] = 8;
            if (gi_build_table(lens, 288, GI_LL_BITS, LL, GI_LL_SIZE, 0, sub_bits, codes) < 0) return -6;
            for (i = 0; i < 32; ++i) lens[i] = 5;
            if (gi_build_table(lens, 32, GI_D_BITS, DD, GI_D_SIZE, 1, sub_bits, codes) < 0) return -6;
        } else {  // dynamic codes
            B.need(14);
            const int hlit = (int)B.peek(5) + 257;
            B.drop(5);
            const int hdist = (int)B.peek(5) + 1;
            B.drop(5);
            const int hclen = (int)B.peek(4) + 4;
            B.drop(4);
            if (hlit > 286 || hdist > 30) return -7;
            uint8_t cl[19];
            for (int i = 0; i < 19; ++i) cl[i] = 0;
            for (int i = 0; i < hclen; ++i) {
                B.need(3);
                cl[c_cl_order[i]] = (uint8_t)B.peek(3);
                B.drop(3);
            }
            uint32_t* cltab = DD;  // 128 entries, free until the distance table is built
            if (gi_build_table(cl, 19, 7, cltab, 1 << 7, 2, sub_bits, codes) < 0) return -8;
            int n = 0;
            const int total = hlit + hdist;
            while (n < total) {
                B.need(14);
                const uint32_t e = cltab[B.peek(7)];
                if (((e >> 5) & 7) != GK_LIT) return -9;
                B.drop((int)(e & 31));
                const int sym = (int)(e >> 16);
                if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
                int rep, val = 0;
                if (sym == 16) {
                    if (n == 0) return -10;
                    val = lens[n - 1];
                    rep = 3 + (int)B.peek(2); B.drop(2);
                } else if (sym == 17) {
                    rep = 3 + (int)B.peek(3); B.drop(3);
                } else {
                    rep = 11 + (int)B.peek(7); B.drop(7);
                }
                if (n + rep > total) return -11;
                for (int i = 0; i < rep; ++i) lens[n + i] = (uint8_t)val;
                n += rep;
            }
            if (lens[256] == 0) return -12;
            if (gi_build_table(lens, hlit, GI_LL_BITS, LL, GI_LL_SIZE, 0, sub_bits, codes) < 0) return -13;
            if (gi_build_table(lens + hlit, hdist, GI_D_BITS, DD, GI_D_SIZE, 1, sub_bits, codes) < 0) return -14;
        }
        // ---- the block's symbols
        while (true) {
            B.need(32);  // a literal / length code (15) + its extra bits (5) fit
            uint32_t e = LL[B.peek(GI_LL_BITS)];
            int kind = (int)((e >> 5) & 7);
            if (kind == GK_SUB) {
                B.drop(GI_LL_BITS);
                e = LL[(e >> 16) + B.peek((int)((e >> 8) & 31))];
                kind = (int)((e >> 5) & 7);
            }
            B.drop((int)(e & 31));
            if (kind == GK_LIT) {
                if (op >= out_len) return -15;
                out[op++] = (uint8_t)(e >> 16);
                continue;
            }
            if (kind == GK_EOB) break;
            if (kind != GK_LEN) return -16;
            const int xl = (int)((e >> 8) & 31);
            const unsigned len = (e >> 16) + B.peek(xl);
            B.drop(xl);
            B.need(32);  // a distance code (15) + its extra bits (13)
            uint32_t de = DD[B.peek(GI_D_BITS)];
            if (((de >> 5) & 7) == GK_SUB) {
                B.drop(GI_D_BITS);
                de = DD[(de >> 16) + B.peek((int)((de >> 8) & 31))];
            }
            if (((de >> 5) & 7) != GK_LEN) return -17;
            B.drop((int)(de & 31));
            const int xd = (int)((de >> 8) & 31);
            const unsigned dist = (de >> 16) + B.peek(xd);
            B.drop(xd);
            if (dist > op || len > out_len - op) return -18;
            const uint8_t* src = out + op - dist;
            uint8_t* dst = out + op;
            for (unsigned i = 0; i < len; ++i) dst[i] = src[i];
            op += len;
        }
    } while (!last);
    if (op != out_len) return -19;
    if (B.over || B.byte_pos() > in + in_len) return -20;
    return 0;
}

struct GiBlock {
    unsigned long long c_off;  // payload offset in the compressed buffer
    unsigned long long u_off;  // destination offset in the inflated buffer
    uint32_t c_len, u_len;
};

// persistent: warps pull blocks from an atomic queue; lane 0 decodes
__global__ void __launch_bounds__(GI_WARPS * 32, 1) k_gi_inflate(const uint8_t* __restrict__ comp, const GiBlock* __restrict__ blk,
                                                                 int n_blocks, uint8_t* out, int* __restrict__ status,
                                                                 unsigned* queue) {
    extern __shared__ __align__(16) uint32_t gi_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane != 0) return;  // the decode chain is serial: one thread per warp, no divergence inside the warp
    uint32_t* tab = gi_smem + (size_t)wid * GI_WARP_WORDS;
    while (true) {
        const int b = (int)atomicAdd(queue, 1u);
        if (b >= n_blocks) break;
        const GiBlock k = blk[b];
        status[b] = k.u_len ? gi_inflate_block(comp + k.c_off, k.c_len, out + k.u_off, k.u_len, tab) : 0;
    }
}

#endif
