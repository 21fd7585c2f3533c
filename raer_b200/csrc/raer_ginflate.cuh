// raer_ginflate.cuh -- raw DEFLATE (RFC 1951) on the device, for BGZF blocks: one warp per block.
//
// A BGZF block is an independent DEFLATE stream of at most 64 KiB of output whose compressed payload and exact output
// size are known before decoding starts (BSIZE / ISIZE), so thousands of blocks decode side by side. Huffman decoding
// itself is a serial chain of table lookups: lane 0 of the warp walks the symbols and turns them into tokens (a literal,
// or a match of `len` bytes `dist` bytes back), GI_TOK at a time, in the warp's slice of shared memory. The whole warp
// then writes the bytes of the batch: a prefix sum over the token lengths gives every token its place, every lane takes
// output bytes lane, lane + 32, ..., finds the token a byte belongs to and follows match references (inside the batch
// through the token table, before the batch through one L2 load) until it reaches a literal or a byte that is already
// in memory. So the copy loop of a match -- a chain of dependent loads from memory the same thread has just written --
// is replaced by independent loads, 32 at a time, and the stores of a batch coalesce.
// The tables of the block live in shared memory as well (10-bit primary table for literals / lengths, 8-bit for
// distances, second-level tables behind them); the compressed bytes are read through a 64-bit bit buffer refilled 32 bits
// at a time from aligned words, one word ahead of its use. Same table layout and the same acceptance rules as the host
// decoder (raer_inflate.cpp); a block the decoder refuses is reported through `status` and the host decides what to do.
#ifndef RAER_GINFLATE_CUH
#define RAER_GINFLATE_CUH

#define GI_LL_BITS 10
#define GI_D_BITS 8
#define GI_LL_SIZE ((1 << GI_LL_BITS) + 320)  // zlib's ENOUGH_LENS bound for a 10-bit root: 1332 entries
#define GI_D_SIZE ((1 << GI_D_BITS) + 512)
#define GI_TOK 64                               // tokens per batch
#define GI_CTRL 4                               // control words lane 0 hands to the warp
#define GI_DSCR 80                              // scratch of the distance-table build (256 prefix bytes + 30 codes)
#define GI_WARP_WORDS (GI_LL_SIZE + GI_D_SIZE + GI_TOK + GI_TOK + 4 + GI_CTRL + GI_DSCR)  // shared-memory words per warp
#define GI_WARPS 24                             // warps (= blocks in flight) per CTA, one CTA per SM

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

// Canonical Huffman decode table, built by the whole warp (every lane calls). what: 0 literal/length alphabet, 1 distance
// alphabet, 2 code-length alphabet. lens / sub_bits / codes live in shared memory. Lane 0 does what is serial by nature
// (code counts, first codes, the code of every symbol: a few instructions per symbol); clearing the table, laying out the
// second-level tables and replicating the entries -- nine tenths of a one-thread build -- are spread over the lanes.
// Returns the number of table entries used, or -1 for an over-subscribed / unusable code or a table that does not fit `cap`.
__device__ int gi_build_table(const uint8_t* lens, int n, int tbits, uint32_t* tab, int cap, int what, uint8_t* sub_bits,
                              uint16_t* codes, int lane) {
    const int primary = 1 << tbits;
    for (int i = lane; i < primary; i += 32) tab[i] = gi_mk(1, GK_BAD, 0, 0);
    for (int i = lane; i < primary / 4; i += 32) reinterpret_cast<uint32_t*>(sub_bits)[i] = 0u;
    __syncwarp();
    int rc = 0;  // lane 0: -1 refused, 1 no symbol at all, 0 go on
    if (lane == 0) {
        int count[16];
        for (int i = 0; i < 16; ++i) count[i] = 0;
        for (int i = 0; i < n; ++i) ++count[lens[i]];
        count[0] = 0;
        int left = 1, used_syms = 0;
        for (int l = 1; l <= 15; ++l) {
            left = (left << 1) - count[l];
            if (left < 0) rc = -1;
            used_syms += count[l];
        }
        if (rc == 0 && used_syms == 0) rc = 1;
        // incomplete codes are legal only with a single code (RFC 1951 / zlib accept one distance code of one bit)
        if (rc == 0 && left > 0 && used_syms != 1) rc = -1;
        if (rc == 0) {
            uint32_t next_code[16];
            uint32_t code = 0;
            for (int l = 1; l <= 15; ++l) { code = (code + count[l - 1]) << 1; next_code[l] = code; }
            for (int sy = 0; sy < n; ++sy) {
                const int l = lens[sy];
                if (!l) continue;
                const uint32_t r = __brev(next_code[l]++) >> (32 - l);
                codes[sy] = (uint16_t)r;
                if (l > tbits) {
                    const uint32_t p = r & (primary - 1);
                    if (l - tbits > sub_bits[p]) sub_bits[p] = (uint8_t)(l - tbits);
                }
            }
        }
    }
    rc = __shfl_sync(0xffffffffu, rc, 0);
    if (rc < 0) return -1;
    if (rc == 1) return primary;
    __syncwarp();
    // second-level tables: lane q lays out the prefixes q * chunk .. in order behind the primary table
    const int chunk = primary / 32;
    int mine = 0;
    for (int p = lane * chunk; p < (lane + 1) * chunk; ++p)
        if (sub_bits[p]) mine += 1 << sub_bits[p];
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const int used = primary + __shfl_sync(0xffffffffu, inc, 31);
    if (used > cap) return -1;
    int base = primary + inc - mine;
    for (int p = lane * chunk; p < (lane + 1) * chunk; ++p)
        if (sub_bits[p]) {
            const int sz = 1 << sub_bits[p];
            tab[p] = gi_mk(tbits, GK_SUB, sub_bits[p], base);
            for (int i = 0; i < sz; ++i) tab[base + i] = gi_mk(1, GK_BAD, 0, 0);
            base += sz;
        }
    __syncwarp();
    for (int sy = lane; sy < n; sy += 32) {
        const int l = lens[sy];
        if (!l) continue;
        uint32_t e;
        if (what == 0) {
            if (sy < 256) e = gi_mk(0, GK_LIT, 0, sy);
            else if (sy == 256) e = gi_mk(0, GK_EOB, 0, 0);
            else if (sy <= 285) e = gi_mk(0, GK_LEN, c_len_extra[sy - 257], c_len_base[sy - 257]);
            else e = gi_mk(0, GK_BAD, 0, 0);
        } else if (what == 1) {
            e = sy < 30 ? gi_mk(0, GK_LEN, c_dist_extra[sy], c_dist_base[sy]) : gi_mk(0, GK_BAD, 0, 0);
        } else {
            e = gi_mk(0, GK_LIT, 0, sy);
        }
        const uint32_t r = codes[sy];
        if (l <= tbits) {
            e |= (uint32_t)l;
            for (uint32_t i = r; i < (uint32_t)primary; i += 1u << l) tab[i] = e;
        } else {
            const uint32_t p = r & (primary - 1);
            const int sb = sub_bits[p];
            const int tb = (int)(tab[p] >> 16);
            e |= (uint32_t)(l - tbits);
            for (uint32_t i = r >> tbits; i < (1u << sb); i += 1u << (l - tbits)) tab[tb + i] = e;
        }
    }
    __syncwarp();
    return used;
}

struct GiBits {  // LSB-first bit reader over aligned 32-bit words of global memory, one word of look-ahead
    const uint32_t* w;      // next word to load (the word after `nxt`)
    const uint32_t* w_end;  // first word that must not be loaded (a few words of slack lie behind the payload)
    unsigned long long bb;
    uint32_t nxt;
    int bc;
    __device__ __forceinline__ uint32_t fetch() {
        const uint32_t v = w < w_end ? __ldg(w) : 0u;  // zeros behind the payload: the caller checks byte_pos() at the end
        ++w;
        return v;
    }
    __device__ __forceinline__ void init(const uint8_t* p, size_t n) {
        const size_t a = (size_t)p & 3;
        w = reinterpret_cast<const uint32_t*>(p - a);
        w_end = reinterpret_cast<const uint32_t*>((((size_t)(p + n)) + 3) & ~(size_t)3) + 2;
        bb = fetch();
        bc = 32;
        nxt = fetch();
        drop((int)a * 8);
    }
    __device__ __forceinline__ void need(int n) {  // n <= 32
        if (bc < n) {
            bb |= (unsigned long long)nxt << bc;
            bc += 32;
            nxt = fetch();
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(int n) { bb >>= n; bc -= n; }
    // first byte of the payload not consumed yet (whole bytes still in the buffer are not consumed)
    __device__ __forceinline__ const uint8_t* byte_pos() const { return reinterpret_cast<const uint8_t*>(w - 1) - (bc >> 3); }
};

// decoder state of one block, in the registers of lane 0
struct GiState {
    GiBits B;
    const uint8_t* in;
    size_t in_len;
    uint32_t op;       // output bytes produced so far, tokens of the current batch included
    uint32_t out_len;
    int last;          // BFINAL of the deflate block being decoded
    int in_symbols;    // tables are built, symbols follow
};

// token: bits 0-8 length (1 = literal), bits 9-16 the literal, bits 17-31 distance - 1
__device__ __forceinline__ uint32_t gi_tok_lit(uint32_t v) { return 1u | (v << 9); }
__device__ __forceinline__ uint32_t gi_tok_match(uint32_t len, uint32_t dist) { return len | ((dist - 1u) << 17); }

// Lane 0, at the start of a deflate block: reads the block header. Returns a negative code, or
//   0  stored block: copied right here (nothing is pending when one starts), S.op moved on
//   1  fixed codes: code lengths written to `lens` (288 literal/length, then 32 distance)
//   2  dynamic codes: *hlit / *hdist set, the 19 code-length code lengths written to `cl`
__device__ int gi_block_header(GiState& S, uint8_t* out, uint8_t* lens, uint8_t* cl, int* hlit, int* hdist) {
    GiBits& B = S.B;
    B.need(3);
    S.last = (int)B.peek(1);
    const int type = (int)(B.peek(3) >> 1);
    B.drop(3);
    if (type == 0) {
        B.drop(B.bc & 7);
        const uint8_t* ip = B.byte_pos();
        if (ip + 4 > S.in + S.in_len) return -2;
        const unsigned len = ip[0] | (ip[1] << 8), nlen = ip[2] | (ip[3] << 8);
        ip += 4;
        if ((len ^ 0xffffu) != nlen) return -3;
        if (ip + len > S.in + S.in_len || S.op + len > S.out_len) return -4;
        for (unsigned i = 0; i < len; ++i) out[S.op + i] = ip[i];
        S.op += len;
        B.init(ip + len, (size_t)((S.in + S.in_len) - (ip + len)));
        return 0;
    }
    if (type == 3) return -5;
    if (type == 1) {
        int i = 0;
        for (; i < 144; ++i) lens[i] = 8;
        for (; i < 256; ++i) lens[i] = 9;
        for (; i < 280; ++i) lens[i] = 7;
        for (; i < 288; ++i) lens[i] = 8;
        for (i = 0; i < 32; ++i) lens[288 + i] = 5;
        *hlit = 288;
        *hdist = 32;
        return 1;
    }
    B.need(14);
    *hlit = (int)B.peek(5) + 257;
    B.drop(5);
    *hdist = (int)B.peek(5) + 1;
    B.drop(5);
    const int hclen = (int)B.peek(4) + 4;
    B.drop(4);
    if (*hlit > 286 || *hdist > 30) return -7;
    for (int i = 0; i < 19; ++i) cl[i] = 0;
    for (int i = 0; i < hclen; ++i) {
        B.need(3);
        cl[c_cl_order[i]] = (uint8_t)B.peek(3);
        B.drop(3);
    }
    return 2;
}

// Lane 0: the code lengths of a dynamic block, through the code-length table
__device__ int gi_code_lengths(GiState& S, const uint32_t* cltab, uint8_t* lens, int total) {
    GiBits& B = S.B;
    int k = 0;
    while (k < total) {
        B.need(14);
        const uint32_t e = cltab[B.peek(7)];
        if (((e >> 5) & 7) != GK_LIT) return -9;
        B.drop((int)(e & 31));
        const int sym = (int)(e >> 16);
        if (sym < 16) { lens[k++] = (uint8_t)sym; continue; }
        int rep, val = 0;
        if (sym == 16) {
            if (k == 0) return -10;
            val = lens[k - 1];
            rep = 3 + (int)B.peek(2); B.drop(2);
        } else if (sym == 17) {
            rep = 3 + (int)B.peek(3); B.drop(3);
        } else {
            rep = 11 + (int)B.peek(7); B.drop(7);
        }
        if (k + rep > total) return -11;
        for (int i = 0; i < rep; ++i) lens[k + i] = (uint8_t)val;
        k += rep;
    }
    return 0;
}

// Lane 0: decodes symbols until GI_TOK tokens are waiting, the deflate block ends or something is wrong.
// Returns 0 = more to come, 1 = stream finished, negative = refused.
__device__ int gi_decode_batch(GiState& S, uint32_t* tab, uint32_t* tok, int* n_tok) {
    uint32_t* const LL = tab;
    uint32_t* const DD = tab + GI_LL_SIZE;
    GiBits& B = S.B;
    int n = 0;
    uint32_t op = S.op;
    int rc = 0;
    while (n < GI_TOK) {
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
            tok[n++] = gi_tok_lit(e >> 16);
            ++op;
            continue;
        }
        if (kind == GK_EOB) {
            S.in_symbols = 0;
            rc = S.last ? 1 : 0;
            break;
        }
        if (kind != GK_LEN) { rc = -16; break; }
        const int xl = (int)((e >> 8) & 31);
        const unsigned len = (e >> 16) + B.peek(xl);
        B.drop(xl);
        B.need(32);  // a distance code (15) + its extra bits (13)
        uint32_t de = DD[B.peek(GI_D_BITS)];
        if (((de >> 5) & 7) == GK_SUB) {
            B.drop(GI_D_BITS);
            de = DD[(de >> 16) + B.peek((int)((de >> 8) & 31))];
        }
        if (((de >> 5) & 7) != GK_LEN) { rc = -17; break; }
        B.drop((int)(de & 31));
        const int xd = (int)((de >> 8) & 31);
        const unsigned dist = (de >> 16) + B.peek(xd);
        B.drop(xd);
        if (dist > op || len > S.out_len - op) { rc = -18; break; }
        tok[n++] = gi_tok_match(len, dist);
        op += len;
    }
    if (rc >= 0 && op > S.out_len) rc = -15;
    S.op = op;
    *n_tok = rc < 0 ? 0 : n;
    return rc;
}

struct GiBlock {
    unsigned long long c_off;  // payload offset in the compressed buffer
    unsigned long long u_off;  // destination offset in the inflated buffer
    uint32_t c_len, u_len;
    uint32_t crc;              // CRC-32 of the inflated bytes (BGZF footer)
    uint32_t check_crc;        // 0: do not verify
};

// ---- CRC-32 (the gzip polynomial, reflected 0xEDB88320) of a block's output by the whole warp. The block is cut into
// chunks of 4 KB, a lane takes 128 consecutive bytes of every chunk (one cache line) and runs the byte-wise table method
// on them; pieces are joined with the shift operator of the CRC (multiplication by x^(8n) modulo the polynomial, the
// algorithm behind zlib's crc32_combine): crc(A || B) = shift(crc(A), |B|) ^ crc(B), linear over XOR, so a lane
// accumulates its own pieces Horner-style across chunks and the 32 lanes are joined once per block.
#define GI_CRC_POLY 0xedb88320u
#define GI_CRC_CHUNK 4096u
#define GI_CRC_WORDS (256 + 32)  // shared per CTA: byte table, x^(2^k) table
__device__ __forceinline__ uint32_t gi_multmodp(uint32_t a, uint32_t b) {  // a(x) * b(x) mod p(x), reflected bit order
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ GI_CRC_POLY : b >> 1;
    }
    return p;
}
// x^(8 n) mod p: x2n[k] = x^(2^k) mod p
__device__ __forceinline__ uint32_t gi_xpow8(uint32_t n_bytes, const uint32_t* x2n) {
    uint32_t p = 1u << 31;  // x^0
    int k = 3;
    while (n_bytes) {
        if (n_bytes & 1u) p = gi_multmodp(x2n[k & 31], p);
        n_bytes >>= 1;
        ++k;
    }
    return p;
}
__device__ void gi_crc_tables(uint32_t* crc_smem) {  // whole CTA; the caller synchronises afterwards
    if (threadIdx.x < 256) {
        uint32_t c = threadIdx.x;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ GI_CRC_POLY : c >> 1;
        crc_smem[threadIdx.x] = c;
    }
    if (threadIdx.x == 256) {
        uint32_t p = 1u << 30;  // x^1
        crc_smem[256] = p;
        for (int k = 1; k < 32; ++k) { p = gi_multmodp(p, p); crc_smem[256 + k] = p; }
    }
}
__device__ uint32_t gi_crc32_warp(const uint8_t* p, uint32_t len, const uint32_t* crc_smem, int lane) {
    const uint32_t* T0 = crc_smem;
    const uint32_t* x2n = crc_smem + 256;
    const uint32_t per = GI_CRC_CHUNK / 32;
    const uint32_t n_full = len / GI_CRC_CHUNK, rem = len - n_full * GI_CRC_CHUNK;
    const uint32_t x_chunk = gi_xpow8(GI_CRC_CHUNK, x2n);
    uint32_t acc = 0;  // finalised CRC of this lane's pieces so far, as if they were contiguous in chunk steps
    for (uint32_t c = 0; c < n_full; ++c) {
        const uint8_t* q = p + (size_t)c * GI_CRC_CHUNK + lane * per;
        uint32_t r = 0xffffffffu;
        for (uint32_t i = 0; i < per; ++i) r = T0[(r ^ q[i]) & 0xffu] ^ (r >> 8);
        acc = gi_multmodp(x_chunk, acc) ^ (r ^ 0xffffffffu);
    }
    uint32_t total = 0;
    if (n_full) {  // bytes behind this lane's last full-chunk piece: the rest of that chunk and the partial chunk
        const uint32_t after = rem + GI_CRC_CHUNK - per * (lane + 1);
        total = gi_multmodp(gi_xpow8(after, x2n), acc);
    }
    if (rem) {
        const uint32_t beg = min(lane * per, rem), end = min(lane * per + per, rem);
        if (end > beg) {
            const uint8_t* q = p + (size_t)n_full * GI_CRC_CHUNK;
            uint32_t r = 0xffffffffu;
            for (uint32_t i = beg; i < end; ++i) r = T0[(r ^ q[i]) & 0xffu] ^ (r >> 8);
            total ^= gi_multmodp(gi_xpow8(rem - end, x2n), r ^ 0xffffffffu);
        }
    }
    return __reduce_xor_sync(0xffffffffu, total);
}

#define GI_SMEM_BYTES ((GI_WARPS * GI_WARP_WORDS + GI_CRC_WORDS) * 4)
// persistent: warps pull blocks from an atomic queue; lane 0 decodes, the warp writes
__global__ void __launch_bounds__(GI_WARPS * 32, 1) k_gi_inflate(const uint8_t* __restrict__ comp, const GiBlock* __restrict__ blk,
                                                                 int n_blocks, uint8_t* out, int* __restrict__ status,
                                                                 unsigned* queue) {
    extern __shared__ __align__(16) uint32_t gi_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* const crc_smem = gi_smem + (size_t)GI_WARPS * GI_WARP_WORDS;
    gi_crc_tables(crc_smem);
    __syncthreads();
    uint32_t* tab = gi_smem + (size_t)wid * GI_WARP_WORDS;
    uint32_t* tok = tab + GI_LL_SIZE + GI_D_SIZE;
    uint32_t* pre = tok + GI_TOK;          // GI_TOK + 1 offsets (+ padding)
    volatile uint32_t* ctrl = pre + GI_TOK + 4;  // [0] tokens in the batch, [1] decoder verdict
    while (true) {
        int b = 0;
        if (lane == 0) b = (int)atomicAdd(queue, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= n_blocks) break;
        const GiBlock k = blk[b];
        uint8_t* const dst = out + k.u_off;
        GiState S;
        if (lane == 0) {
            S.in = comp + k.c_off;
            S.in_len = k.c_len;
            S.op = 0;
            S.out_len = k.u_len;
            S.last = 0;
            S.in_symbols = 0;
            S.B.init(S.in, S.in_len);
        }
        int verdict = 0;
        uint32_t base = 0;  // output offset of the batch
        if (k.u_len == 0) verdict = 1;
        // scratch of the table builds: code lengths (and the 19 code-length code lengths behind them) in the token area,
        // which is idle while a block header is read; prefix bits / codes of the literal table in the distance table's
        // space (built afterwards), those of the distance table in a scratch of their own
        uint8_t* const lens = reinterpret_cast<uint8_t*>(tok);
        uint8_t* const cl = lens + 460;
        uint32_t* const DD = tab + GI_LL_SIZE;
        uint32_t* const dscr = const_cast<uint32_t*>(ctrl) + GI_CTRL;
        int in_symbols = 0;  // warp-uniform copy of lane 0's S.in_symbols
        while (verdict == 0) {
            if (!in_symbols) {
                if (lane == 0) {
                    int hlit = 0, hdist = 0;
                    const int rc = gi_block_header(S, dst, lens, cl, &hlit, &hdist);
                    ctrl[0] = (uint32_t)rc;
                    ctrl[1] = (uint32_t)hlit;
                    ctrl[2] = (uint32_t)hdist;
                    ctrl[3] = (uint32_t)S.last;
                }
                __syncwarp();
                int hrc = (int)ctrl[0];
                const int hlit = (int)ctrl[1], hdist = (int)ctrl[2];
                const int last = (int)ctrl[3];
                __syncwarp();
                if (hrc < 0) { verdict = hrc; break; }
                if (hrc == 0) {  // stored block, already copied
                    if (last) verdict = 1;
                    continue;
                }
                if (hrc == 2) {
                    if (gi_build_table(cl, 19, 7, DD, 1 << 7, 2, reinterpret_cast<uint8_t*>(DD + 128),
                                       reinterpret_cast<uint16_t*>(DD + 128 + 32), lane) < 0) { verdict = -8; break; }
                    if (lane == 0) {
                        int rc = gi_code_lengths(S, DD, lens, hlit + hdist);
                        if (rc == 0 && lens[256] == 0) rc = -12;
                        ctrl[0] = (uint32_t)rc;
                    }
                    __syncwarp();
                    hrc = (int)ctrl[0];
                    __syncwarp();
                    if (hrc < 0) { verdict = hrc; break; }
                }
                if (gi_build_table(lens, hlit, GI_LL_BITS, tab, GI_LL_SIZE, 0, reinterpret_cast<uint8_t*>(DD),
                                   reinterpret_cast<uint16_t*>(DD + 256), lane) < 0) { verdict = -13; break; }
                if (gi_build_table(lens + hlit, hdist, GI_D_BITS, DD, GI_D_SIZE, 1, reinterpret_cast<uint8_t*>(dscr),
                                   reinterpret_cast<uint16_t*>(dscr + 64), lane) < 0) { verdict = -14; break; }
                in_symbols = 1;
                if (lane == 0) S.in_symbols = 1;
            }
            if (lane == 0) {
                int n = 0;
                const uint32_t before = S.op;
                const int rc = gi_decode_batch(S, tab, tok, &n);
                ctrl[0] = (uint32_t)n;
                ctrl[1] = (uint32_t)rc;
                ctrl[2] = before;
                ctrl[3] = (uint32_t)S.in_symbols;
            }
            __syncwarp();
            const int n = (int)ctrl[0];
            verdict = (int)ctrl[1];
            base = ctrl[2];
            in_symbols = (int)ctrl[3];
            if (n > 0) {
                // place of every token: lane handles tokens 2 * lane and 2 * lane + 1
                const uint32_t l0 = 2 * lane < n ? (tok[2 * lane] & 511u) : 0u;
                const uint32_t l1 = 2 * lane + 1 < n ? (tok[2 * lane + 1] & 511u) : 0u;
                uint32_t inc = l0 + l1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                pre[2 * lane] = inc - l0 - l1;
                pre[2 * lane + 1] = inc - l1;
                if (lane == 31) pre[GI_TOK] = inc;
                const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
                __syncwarp();
                if (total == (uint32_t)n) {  // literals only
                    for (uint32_t j = lane; j < total; j += 32) dst[base + j] = (uint8_t)(tok[j] >> 9);
                } else {
                    // four bytes per lane in flight: the loads of the bytes that lie before the batch are independent
                    for (uint32_t j0 = lane; j0 < total; j0 += 128) {
                        int src[4];
                        uint32_t val[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            int pos = (int)(j0 + 32 * q);
                            src[q] = 0;
                            val[q] = 0;
                            if ((uint32_t)pos >= total) continue;
                            while (true) {
                                int t = 0;  // last token with pre[t] <= pos
#pragma unroll
                                for (int step = GI_TOK / 2; step > 0; step >>= 1)
                                    if (pre[t + step] <= (uint32_t)pos) t += step;
                                const uint32_t tk = tok[t];
                                if ((tk & 511u) == 1u) { val[q] = (tk >> 9) & 255u; break; }
                                const int dist = (int)(tk >> 17) + 1;
                                int within = pos - (int)pre[t];
                                if (within >= dist) within %= dist;
                                pos = (int)pre[t] - dist + within;
                                if (pos < 0) { src[q] = pos; break; }  // written by an earlier batch
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (src[q] < 0) val[q] = __ldcg(dst + (int)base + src[q]);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (j0 + 32 * q < total) dst[base + j0 + 32 * q] = (uint8_t)val[q];
                    }
                }
            }
            __syncwarp();  // the bytes of this batch are in memory before the next batch refers to them
        }
        int rc = verdict == 1 ? 0 : verdict;
        if (lane == 0 && rc == 0 && k.u_len) {
            if (S.op != S.out_len) rc = -19;
            else if (S.B.byte_pos() > S.in + S.in_len) rc = -20;
        }
        rc = __shfl_sync(0xffffffffu, rc, 0);
        if (rc == 0 && k.check_crc) {  // gzip members carry the CRC-32 of their content: verified like htslib does
            __syncwarp();
            if (gi_crc32_warp(dst, k.u_len, crc_smem, lane) != k.crc) rc = -21;
        }
        if (lane == 0) status[b] = rc;
    }
}

#endif
