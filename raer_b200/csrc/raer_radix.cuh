// raer_radix.cuh -- LSD radix sort of 64-bit keys with 32-bit values (8-bit digits, stable block scatter), shared by the
// single-cell tuple sort (raer_sc.cu) and the read-name pairing of the device ingest (raer_gingest.cu).
#ifndef RAER_RADIX_CUH
#define RAER_RADIX_CUH
// ------------------------------------------------------------------------------------- radix sort
#define RS_BLOCK 256
#define RS_ITEMS 16  // keys per thread per block chunk

static __global__ void k_rs_hist(const unsigned long long* __restrict__ keys, long long n, int shift, unsigned* __restrict__ hist) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    long long base = (long long)blockIdx.x * RS_BLOCK * RS_ITEMS;
    for (int i = 0; i < RS_ITEMS; ++i) {
        long long j = base + (long long)i * RS_BLOCK + threadIdx.x;
        if (j < n) atomicAdd(&h[(keys[j] >> shift) & 0xff], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];  // digit-major
}

static __device__ __forceinline__ unsigned block_incl_scan_256(unsigned v, unsigned* part) {
    part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        unsigned t = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += t;
        __syncthreads();
    }
    return part[threadIdx.x];
}

// hist is digit-major [256][nblk]. Block d turns row d into an exclusive scan and writes the row total;
// k_rs_scan_tot then turns the 256 totals into digit bases.
static __global__ void k_rs_scan_digit(unsigned* hist, int nblk, unsigned* tot) {
    __shared__ unsigned part[256];
    __shared__ unsigned carry;
    unsigned* row = hist + (size_t)blockIdx.x * nblk;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 256) {
        int i = base + threadIdx.x;
        unsigned v = i < nblk ? row[i] : 0;
        unsigned inc = block_incl_scan_256(v, part);
        if (i < nblk) row[i] = carry + inc - v;
        __syncthreads();
        if (threadIdx.x == 255) carry += inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) tot[blockIdx.x] = carry;
}
static __global__ void k_rs_scan_tot(unsigned* tot) {
    __shared__ unsigned part[256];
    unsigned v = tot[threadIdx.x];
    unsigned inc = block_incl_scan_256(v, part);
    tot[threadIdx.x] = inc - v;
}

static __global__ void k_rs_scatter(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ vals, long long n,
                             int shift, const unsigned* __restrict__ offs, const unsigned* __restrict__ dbase,
                             unsigned long long* __restrict__ okeys, unsigned* __restrict__ ovals) {
    __shared__ unsigned wcnt[RS_BLOCK / 32][256];
    __shared__ unsigned run[256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    run[threadIdx.x] = dbase[threadIdx.x] + offs[(size_t)threadIdx.x * gridDim.x + blockIdx.x];
    long long base = (long long)blockIdx.x * RS_BLOCK * RS_ITEMS;
    for (int i = 0; i < RS_ITEMS; ++i) {
        for (int d = threadIdx.x; d < (RS_BLOCK / 32) * 256; d += RS_BLOCK) (&wcnt[0][0])[d] = 0;
        __syncthreads();
        long long j = base + (long long)i * RS_BLOCK + threadIdx.x;
        bool valid = j < n;
        unsigned long long key = valid ? keys[j] : 0ull;
        unsigned dig = valid ? (unsigned)((key >> shift) & 0xff) : 0xffffffffu;
        unsigned peers = __match_any_sync(0xffffffffu, dig);
        unsigned rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) wcnt[w][dig] = __popc(peers);
        __syncthreads();
        if (valid) {
            unsigned before = 0;
            for (int ww = 0; ww < w; ++ww) before += wcnt[ww][dig];
            unsigned dst = run[dig] + before + rank;
            okeys[dst] = key;
            ovals[dst] = vals[j];
        }
        __syncthreads();
        {
            unsigned tot = 0;
            for (int ww = 0; ww < RS_BLOCK / 32; ++ww) tot += wcnt[ww][threadIdx.x];
            run[threadIdx.x] += tot;
        }
        __syncthreads();
    }
}


// sorts n (key, value) pairs by the low key_bits of the key; returns the buffers that hold the result through k_in / v_in.
// hist: nblk * 256 + 256 words of scratch, nblk = ceil(n / (RS_BLOCK * RS_ITEMS)). Returns the number of launches.
static inline int rs_sort(unsigned long long*& k_in, unsigned*& v_in, unsigned long long*& k_out, unsigned*& v_out, long long n,
                          int key_bits, unsigned* hist, cudaStream_t st) {
    if (n <= 0) return 0;
    const unsigned nblk = (unsigned)((n + (long long)RS_BLOCK * RS_ITEMS - 1) / ((long long)RS_BLOCK * RS_ITEMS));
    unsigned* d_dtot = hist + (size_t)nblk * 256;
    int nl = 0;
    for (int shift = 0; shift < key_bits; shift += 8) {
        k_rs_hist<<<nblk, RS_BLOCK, 0, st>>>(k_in, n, shift, hist);
        k_rs_scan_digit<<<256, 256, 0, st>>>(hist, (int)nblk, d_dtot);
        k_rs_scan_tot<<<1, 256, 0, st>>>(d_dtot);
        k_rs_scatter<<<nblk, RS_BLOCK, 0, st>>>(k_in, v_in, n, shift, hist, d_dtot, k_out, v_out);
        unsigned long long* tk = k_in; k_in = k_out; k_out = tk;
        unsigned* tv = v_in; v_in = v_out; v_out = tv;
        nl += 4;
    }
    return nl;
}
#endif
