"""The mate-overlap kernel (k_mate_overlap: single-block fast path, cursor fast-forward, 8 qualities per
step) against the oracle's restatement of htslib's tweak_overlap_quality, on the CPU: the device code is
taken verbatim from raer_kernels.cu, compiled for the host with the few CUDA intrinsics it uses supplied
as plain C++, and run next to oracle/plp_sim.c on random CIGAR pairs (both overlap rules). No GPU needed:
this pins the kernel's control flow and byte arithmetic; the -m gpu suites pin the real thing."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HOST_PRELUDE = r'''
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
using std::min; using std::max;
#include "raer_b200.h"
static inline int cg_op(uint32_t c) { return (int)(c & 0xf); }
static inline int cg_len(uint32_t c) { return (int)(c >> 4); }
enum { OP_M = 0, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X };
static inline int seq_code(const uint8_t* seq, int q) { return (seq[q >> 1] >> ((~q & 1) << 2)) & 0xf; }
static inline uint32_t nib_swap(uint32_t w) { return ((w & 0x0f0f0f0fu) << 4) | ((w >> 4) & 0x0f0f0f0fu); }
static inline uint32_t nib_nonzero(uint32_t x) { uint32_t t = x | (x >> 2); t |= t >> 1; return t & 0x11111111u; }
static inline int __ffs(uint32_t x) { return __builtin_ffs(x); }
static inline uint32_t __vminu2(uint32_t a, uint32_t b) { return min(a & 0xffffu, b & 0xffffu) | (min(a >> 16, b >> 16) << 16); }
'''

HOST_MAIN = r'''
extern "C" int oracle_tweak(int apos, int an, const uint32_t* acig, int alq, const uint8_t* aseq, uint8_t* aq, const char* qname,
                            int bpos, int bn, const uint32_t* bcig, int blq, const uint8_t* bseq, uint8_t* bq, int mode);
extern "C" int oracle_keep_a(const char* qname);

static std::vector<uint32_t> randcig(int& lq, int& rlen) {
    std::vector<uint32_t> c; lq = 0; rlen = 0;
    auto add = [&](int op, int len) { c.push_back((len << 4) | op); if (op == OP_M || op == OP_I || op == OP_S || op == OP_EQ || op == OP_X) lq += len;
                                      if (op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X) rlen += len; };
    if (rand() % 4 == 0) add(OP_H, 3);
    if (rand() % 4 == 0) add(OP_S, 1 + rand() % 9);
    int nb = (rand() % 2) ? 1 : 1 + rand() % 4;
    for (int i = 0; i < nb; i++) {
        int mop = rand() % 10 == 0 ? (rand() % 2 ? OP_EQ : OP_X) : OP_M;
        add(mop, 1 + rand() % 60);
        if (i + 1 < nb) { int r = rand() % 5; if (r == 0) add(OP_I, 1 + rand() % 3); else if (r == 1) add(OP_D, 1 + rand() % 4); else if (r == 2) add(OP_N, 5 + rand() % 60); else if (r == 3) add(OP_P, 1); else add(OP_D, 1 + rand() % 3); }
    }
    if (rand() % 4 == 0) add(OP_S, 1 + rand() % 9);
    return c;
}
int main(int argc, char** argv) {
    const int n_cases = argc > 1 ? atoi(argv[1]) : 200000;
    srand(12345);
    long long bad = 0;
    for (int it = 0; it < n_cases; ++it) {
        int lqa, rla, lqb, rlb;
        auto ca = randcig(lqa, rla), cb = randcig(lqb, rlb);
        raer_read_hdr h[2]; memset(h, 0, sizeof(h));
        std::vector<uint32_t> cig(ca); cig.insert(cig.end(), cb.begin(), cb.end());
        const int apos = 100, bpos = 100 + rand() % (rla + 3);
        char qname[32]; snprintf(qname, sizeof qname, "read%d", rand());
        h[0].pos = apos; h[0].end = apos + rla; h[0].l_qseq = lqa; h[0].n_cigar = ca.size(); h[0].cigar_off = 0; h[0].sq_off = 0; h[0].mate = -1;
        h[1].pos = bpos; h[1].end = bpos + rlb; h[1].l_qseq = lqb; h[1].n_cigar = cb.size(); h[1].cigar_off = ca.size(); h[1].sq_off = 256; h[1].mate = 0;
        h[1].bits = oracle_keep_a(qname) ? RAER_RB_KEEP_A : 0;
        std::vector<uint8_t> seq(1024), q1(2048), q2;
        for (auto& x : seq) x = (uint8_t)(((1 << (rand() % 4 ? 0 : rand() % 4)) << 4) | (1 << (rand() % 4 ? 0 : rand() % 4)));
        for (auto& x : q1) x = (uint8_t)(rand() % 3 == 0 ? rand() % 210 : 37);
        q2 = q1;
        int32_t pb = 1;
        const int mode = 1 + (it & 1);
        k_mate_overlap(h, cig.data(), seq.data(), q1.data(), &pb, 1, mode, nullptr, nullptr, 0);
        oracle_tweak(apos, (int)ca.size(), ca.data(), lqa, seq.data(), q2.data(), qname,
                     bpos, (int)cb.size(), cb.data(), lqb, seq.data() + 256, q2.data() + 512, mode);
        // the kernel may rewrite the padding of a read's last 8-byte quality word with its own value: compare the bases only
        bool diff = memcmp(q1.data(), q2.data(), lqa) != 0 || memcmp(q1.data() + 512, q2.data() + 512, lqb) != 0;
        if (diff) {
            if (bad < 5) {
                printf("DIFF case %d mode %d apos %d bpos %d A:", it, mode, apos, bpos);
                for (auto w : ca) printf("%d%c", w >> 4, "MIDNSHP=X"[w & 15]);
                printf(" B:");
                for (auto w : cb) printf("%d%c", w >> 4, "MIDNSHP=X"[w & 15]);
                printf("\n");
            }
            bad++;
        }
    }
    printf("cases=%d bad=%lld\n", n_cases, bad);
    return bad != 0;
}
'''

ORACLE_SHIM = r'''
#include "plp_sim.c"
int oracle_keep_a(const char* qname) { return (int)(wang(x31(qname)) & 1); }
int oracle_tweak(int apos, int an, const uint32_t* acig, int alq, const uint8_t* aseq, uint8_t* aq, const char* qname,
                 int bpos, int bn, const uint32_t* bcig, int blq, const uint8_t* bseq, uint8_t* bq, int mode) {
    rec_t a, b;
    memset(&a, 0, sizeof a); memset(&b, 0, sizeof b);
    a.pos = apos; a.n_cigar = an; a.cigar = (uint32_t*)acig; a.l_qseq = alq; a.seq = (uint8_t*)aseq; a.qual = aq; a.qname = (char*)qname;
    b.pos = bpos; b.n_cigar = bn; b.cigar = (uint32_t*)bcig; b.l_qseq = blq; b.seq = (uint8_t*)bseq; b.qual = bq; b.qname = (char*)qname;
    return tweak_overlap_quality(&a, &b, mode);
}
'''


def _host_kernel_source():
    s = open(os.path.join(ROOT, "raer_b200", "csrc", "raer_kernels.cu")).read()
    i0 = s.index("struct CigCur {")
    i1 = s.index("// ---------------------------------------------------------------------------------------------\n// tile read ranges")
    body = s[i0:i1]
    # the warp-cooperative kernel for single-block pairs needs shuffles: leave it to the -m gpu suites; the
    # thread-per-pair kernel below still takes such pairs through its own single-block path when given no list
    j0 = body.index("// Pairs whose mates are both one aligned block: the overlap is one run, known up front.")
    j1 = body.index("// Pairs with a CIGAR to walk")
    body = body[:j0] + body[j1:]
    for a, b in (("__device__ ", ""), ("__global__ ", ""), ("__restrict__", ""), ("__forceinline__ ", "inline ")):
        body = body.replace(a, b)
    launch = ("long long n_pairs, int mode, const int*  list, const int*  n_list) {\n"
              "    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;\n")
    assert launch in body, "k_mate_overlap prologue changed: update the host harness"
    return body.replace(launch, "long long n_pairs, int mode, const int* list, const int* n_list, long long t) {\n")


def test_overlap_kernel_matches_oracle_on_random_cigars(tmp_path):
    d = str(tmp_path)
    open(os.path.join(d, "kernel_host.cpp"), "w").write(HOST_PRELUDE + _host_kernel_source() + HOST_MAIN)
    open(os.path.join(d, "oracle_shim.c"), "w").write(ORACLE_SHIM)
    inc = ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle")]
    subprocess.run(["gcc", "-O1", "-std=gnu11", "-c", "-o", os.path.join(d, "oracle_shim.o"), os.path.join(d, "oracle_shim.c")] + inc,
                   check=True)
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", os.path.join(d, "harness"), os.path.join(d, "kernel_host.cpp"),
                    os.path.join(d, "oracle_shim.o"), os.path.join(ROOT, "oracle", "liboracle.so"),
                    "-Wl,-rpath," + os.path.join(ROOT, "oracle")] + inc, check=True)
    p = subprocess.run([os.path.join(d, "harness"), "100000"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "bad=0" in p.stdout
