#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <chrono>
#include <zlib.h>
#include "raer_inflate.h"

static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t>& src, int level, int strategy) {
    z_stream zs; memset(&zs, 0, sizeof(zs));
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
    std::vector<uint8_t> out(deflateBound(&zs, src.size()) + 64);
    zs.next_in = (Bytef*)src.data(); zs.avail_in = src.size();
    zs.next_out = out.data(); zs.avail_out = out.size();
    int r = deflate(&zs, Z_FINISH);
    if (r != Z_STREAM_END) { fprintf(stderr, "deflate failed\n"); exit(1); }
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return out;
}
static unsigned rs = 12345;
static unsigned rnd() { rs = rs * 1103515245u + 12345u; return rs >> 8; }

int main(int argc, char** argv) {
    int bad = 0, n = 0, fallback = 0;
    for (int iter = 0; iter < 1500; ++iter) {
        int kind = rnd() % 6;
        size_t len = (iter % 50 == 0) ? 65280 : rnd() % 65281;
        if (iter % 97 == 0) len = rnd() % 20;
        std::vector<uint8_t> src(len);
        for (size_t i = 0; i < len; ++i) {
            switch (kind) {
                case 0: src[i] = rnd() & 0xff; break;                              // incompressible
                case 1: src[i] = "ACGT"[rnd() & 3]; break;                         // 2 bits of entropy
                case 2: src[i] = (uint8_t)(i % 7 == 0 ? rnd() & 0xff : 'I'); break;  // long runs
                case 3: src[i] = i > 100 && (rnd() % 10) ? src[i - 1 - rnd() % 100] : (rnd() & 0x3f); break;  // matches
                case 4: src[i] = 0; break;
                default: src[i] = (uint8_t)((i * 31) ^ (i >> 5)); break;
            }
        }
        int level = rnd() % 10;
        int strat = (int[]){Z_DEFAULT_STRATEGY, Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE, Z_FILTERED}[rnd() % 6];
        std::vector<uint8_t> c = deflate_raw(src, level, strat);
        size_t clen = c.size();
        c.resize(clen + 16, 0xAB);  // readable slack
        std::vector<uint8_t> out(len + 32, 0xCD);
        int rc = raer::inflate_raw(c.data(), clen, out.data() + 16, len);
        ++n;
        if (rc != 0) { ++fallback; fprintf(stderr, "iter %d kind %d level %d strat %d len %zu: decoder refused\n", iter, kind, level, strat, len); continue; }
        if (memcmp(out.data() + 16, src.data(), len) != 0) { ++bad; fprintf(stderr, "iter %d: MISMATCH\n", iter); }
        for (int k = 0; k < 16; ++k) if (out[k] != 0xCD || out[16 + len + k] != 0xCD) { ++bad; fprintf(stderr, "iter %d: wrote outside\n", iter); break; }
        // truncated / corrupted input must not crash (result is allowed to be an error or garbage, never an overrun)
        if (clen > 4) {
            std::vector<uint8_t> c2 = c;
            c2[rnd() % clen] ^= 1 << (rnd() % 8);
            std::vector<uint8_t> o2(len + 32, 0xCD);
            raer::inflate_raw(c2.data(), clen, o2.data() + 16, len);
            for (int k = 0; k < 16; ++k) if (o2[k] != 0xCD || o2[16 + len + k] != 0xCD) { ++bad; fprintf(stderr, "iter %d: corrupted input wrote outside\n", iter); break; }
        }
    }
    printf("%d streams, %d mismatches, %d refused\n", n, bad, fallback);
    // speed on BAM-like data
    std::vector<uint8_t> src(65280);
    for (size_t i = 0; i < src.size(); ++i) src[i] = (i % 230) < 60 ? (uint8_t)(rnd() & 0xff) : ((i % 230) < 110 ? (uint8_t)(rnd() % 16 * 17) : (rnd() % 10 ? 'F' : (uint8_t)(35 + rnd() % 12)));
    std::vector<uint8_t> c = deflate_raw(src, 6, Z_DEFAULT_STRATEGY);
    size_t clen = c.size(); c.resize(clen + 16);
    std::vector<uint8_t> out(src.size());
    auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < 2000; ++k) raer::inflate_raw(c.data(), clen, out.data(), out.size());
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("own inflate: %.0f MB/s (ratio %.2f)\n", 2000.0 * src.size() / dt / 1e6, (double)src.size() / clen);
    t0 = std::chrono::steady_clock::now();
    z_stream zs; memset(&zs, 0, sizeof(zs)); inflateInit2(&zs, -15);
    for (int k = 0; k < 2000; ++k) { inflateReset(&zs); zs.next_in = c.data(); zs.avail_in = clen; zs.next_out = out.data(); zs.avail_out = out.size(); inflate(&zs, Z_FINISH); }
    dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("zlib inflate: %.0f MB/s\n", 2000.0 * src.size() / dt / 1e6);
    return bad ? 1 : 0;
}
