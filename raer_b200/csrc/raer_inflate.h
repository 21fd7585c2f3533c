// Raw DEFLATE decoder for BGZF blocks (RFC 1951), written for this ingest: the whole compressed payload and the
// exact output size of a block are known up front, so there is no streaming state, the bit buffer is refilled with
// one unaligned 64-bit load, literals and matches are decoded from one 11-bit table lookup each and matches are
// copied eight bytes at a time. Returns 0 on success; any inconsistency returns -1 and the caller falls back to zlib.
// Never writes outside [out, out + out_len) (neighbouring blocks are inflated by other threads) and never reads
// outside [in, in + in_len + 16): the caller guarantees 16 readable bytes behind the payload (in a BGZF file the
// 8-byte footer and the next block's header; the last block of a buffer goes to zlib).
#ifndef RAER_INFLATE_H
#define RAER_INFLATE_H
#include <stdint.h>
#include <stddef.h>

namespace raer {
int inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);
}
#endif
