"""Tiny pure-Python BAM/BAI reader-writer for test fixtures (not used by the product).

Plays the role of Rsamtools::filterBam / asBam / indexBam in the reference's tests
(tests/testthat/test_pileup_sites.R:205-240, test_pileup_cells.R:287-361).
"""
from __future__ import annotations

import struct
import zlib
from typing import Callable, List, Tuple

_NT16 = "=ACMGRSVTWYHKDBN"
_NT16_CODE = {c: i for i, c in enumerate(_NT16)}
_CIGAR_OPS = "MIDNSHP=XB"


def _bgzf_blocks(path):
    with open(path, "rb") as fh:
        data = fh.read()
    o = 0
    out = bytearray()
    while o < len(data):
        xlen = struct.unpack("<H", data[o + 10:o + 12])[0]
        bsize = None
        e = o + 12
        while e < o + 12 + xlen:
            slen = struct.unpack("<H", data[e + 2:e + 4])[0]
            if data[e:e + 2] == b"BC":
                bsize = struct.unpack("<H", data[e + 4:e + 6])[0]
            e += 4 + slen
        cdata = data[o + 12 + xlen:o + bsize + 1 - 8]
        out += zlib.decompress(cdata, -15)
        o += bsize + 1
    return bytes(out)


class Rec:
    __slots__ = ("raw", "tid", "pos", "mapq", "flag", "l_seq", "n_cigar", "l_qname", "mtid", "mpos", "isize")

    def __init__(self, raw: bytes):
        self.raw = raw
        (self.tid, self.pos, self.l_qname, self.mapq, _bin, self.n_cigar, self.flag, self.l_seq, self.mtid, self.mpos,
         self.isize) = struct.unpack("<iiBBHHHIiii", raw[:32])

    @property
    def qname(self) -> str:
        return self.raw[32:32 + self.l_qname - 1].decode()

    @property
    def cigar(self) -> List[Tuple[int, int]]:
        o = 32 + self.l_qname
        return [(v & 0xF, v >> 4) for v in struct.unpack(f"<{self.n_cigar}I", self.raw[o:o + 4 * self.n_cigar])]

    @property
    def qual(self) -> bytes:
        o = 32 + self.l_qname + 4 * self.n_cigar + (self.l_seq + 1) // 2
        return self.raw[o:o + self.l_seq]

    @property
    def seq(self) -> str:
        o = 32 + self.l_qname + 4 * self.n_cigar
        b = self.raw[o:o + (self.l_seq + 1) // 2]
        return "".join(_NT16[(b[i >> 1] >> (4 if i % 2 == 0 else 0)) & 0xF] for i in range(self.l_seq))

    def ref_len(self) -> int:
        return sum(l for op, l in self.cigar if op in (0, 2, 3, 7, 8))

    def tag_z(self, tag: str):
        o = 32 + self.l_qname + 4 * self.n_cigar + (self.l_seq + 1) // 2 + self.l_seq
        raw = self.raw
        sizes = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4, "d": 8}
        while o + 3 <= len(raw):
            t = raw[o:o + 2].decode()
            ty = chr(raw[o + 2])
            o += 3
            if ty in ("Z", "H"):
                e = raw.index(b"\0", o)
                if t == tag:
                    return raw[o:e].decode()
                o = e + 1
            elif ty == "B":
                sub = chr(raw[o])
                cnt = struct.unpack("<I", raw[o + 1:o + 5])[0]
                o += 5 + sizes[sub] * cnt
            else:
                o += sizes[ty]
        return None


def read_bam(path):
    """returns (header_text, [(name, len)], [Rec])"""
    d = _bgzf_blocks(path)
    assert d[:4] == b"BAM\x01"
    l_text = struct.unpack("<i", d[4:8])[0]
    text = d[8:8 + l_text].decode()
    o = 8 + l_text
    n_ref = struct.unpack("<i", d[o:o + 4])[0]
    o += 4
    refs = []
    for _ in range(n_ref):
        ln = struct.unpack("<i", d[o:o + 4])[0]
        name = d[o + 4:o + 4 + ln - 1].decode()
        L = struct.unpack("<i", d[o + 4 + ln:o + 8 + ln])[0]
        refs.append((name, L))
        o += 8 + ln
    recs = []
    while o < len(d):
        bs = struct.unpack("<i", d[o:o + 4])[0]
        recs.append(Rec(d[o + 4:o + 4 + bs]))
        o += 4 + bs
    return text, refs, recs


def _reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def _bgzf_block(payload: bytes) -> bytes:
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    c = co.compress(payload) + co.flush()
    bsize = len(c) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + c +
            struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload)))


_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def write_bam(path, header_text: str, refs, recs: List[Rec], block_payload: int = 60000):
    """Writes <path> and <path>.bai. Records must be coordinate sorted."""
    hdr = bytearray(b"BAM\x01")
    tb = header_text.encode()
    hdr += struct.pack("<i", len(tb)) + tb + struct.pack("<i", len(refs))
    for name, L in refs:
        nb = name.encode() + b"\0"
        hdr += struct.pack("<i", len(nb)) + nb + struct.pack("<i", L)
    out = bytearray()
    out += _bgzf_block(bytes(hdr))
    # per-reference index accumulators
    bins = [dict() for _ in refs]
    lin = [dict() for _ in refs]
    cur = bytearray()
    pending = []  # (rec, offset_in_block)

    def flush():
        nonlocal cur, pending
        if not cur:
            return
        coff = len(out)
        blk = _bgzf_block(bytes(cur))
        out.extend(blk)
        next_coff = len(out)
        for i, (r, uo) in enumerate(pending):
            if r.tid < 0:
                continue
            vbeg = (coff << 16) | uo
            if i + 1 < len(pending):
                vend = (coff << 16) | pending[i + 1][1]
            else:
                vend = next_coff << 16
            end = r.pos + (r.ref_len() or 1)
            b = _reg2bin(r.pos, end)
            ch = bins[r.tid].setdefault(b, [])
            if ch and ch[-1][1] == vbeg:
                ch[-1][1] = vend
            else:
                ch.append([vbeg, vend])
            for w in range(r.pos >> 14, ((end - 1) >> 14) + 1):
                if w not in lin[r.tid] or vbeg < lin[r.tid][w]:
                    lin[r.tid][w] = vbeg
        cur = bytearray()
        pending = []

    for r in recs:
        blob = struct.pack("<i", len(r.raw)) + r.raw
        if len(cur) + len(blob) > block_payload:
            flush()
        pending.append((r, len(cur)))
        cur += blob
    flush()
    out += _EOF
    with open(path, "wb") as fh:
        fh.write(out)
    bai = bytearray(b"BAI\x01" + struct.pack("<i", len(refs)))
    for t in range(len(refs)):
        bai += struct.pack("<i", len(bins[t]))
        for b, chunks in sorted(bins[t].items()):
            bai += struct.pack("<Ii", b, len(chunks))
            for vb, ve in chunks:
                bai += struct.pack("<QQ", vb, ve)
        n_intv = (max(lin[t]) + 1) if lin[t] else 0
        bai += struct.pack("<i", n_intv)
        last = 0
        for w in range(n_intv):
            if w in lin[t]:
                last = lin[t][w]
            bai += struct.pack("<Q", last)
    with open(path + ".bai", "wb") as fh:
        fh.write(bai)
    return path


def write_bam_straddling(path, header_text: str, refs, recs: List[Rec], block_payload: int = 3000):
    """Like write_bam, but the record stream is cut into blocks of exactly block_payload bytes wherever that falls, so
    records straddle BGZF blocks (legal BAM; htsjdk-style writers produce it). Writes <path> and <path>.bai."""
    hdr = bytearray(b"BAM\x01")
    tb = header_text.encode()
    hdr += struct.pack("<i", len(tb)) + tb + struct.pack("<i", len(refs))
    for name, L in refs:
        nb = name.encode() + b"\0"
        hdr += struct.pack("<i", len(nb)) + nb + struct.pack("<i", L)
    out = bytearray(_bgzf_block(bytes(hdr)))
    stream = bytearray()
    starts = []
    for r in recs:
        starts.append(len(stream))
        stream += struct.pack("<i", len(r.raw)) + r.raw
    coffs = []
    for k in range(0, len(stream), block_payload):
        coffs.append(len(out))
        out += _bgzf_block(bytes(stream[k:k + block_payload]))
    coffs.append(len(out))
    out += _EOF

    def voff(x):
        return (coffs[x // block_payload] << 16) | (x % block_payload)

    bins = [dict() for _ in refs]
    lin = [dict() for _ in refs]
    for i, r in enumerate(recs):
        if r.tid < 0:
            continue
        vbeg = voff(starts[i])
        vend = voff(starts[i + 1]) if i + 1 < len(recs) else voff(len(stream))
        end = r.pos + (r.ref_len() or 1)
        ch = bins[r.tid].setdefault(_reg2bin(r.pos, end), [])
        if ch and ch[-1][1] == vbeg:
            ch[-1][1] = vend
        else:
            ch.append([vbeg, vend])
        for w in range(r.pos >> 14, ((end - 1) >> 14) + 1):
            if w not in lin[r.tid] or vbeg < lin[r.tid][w]:
                lin[r.tid][w] = vbeg
    with open(path, "wb") as fh:
        fh.write(out)
    bai = bytearray(b"BAI\x01" + struct.pack("<i", len(refs)))
    for t in range(len(refs)):
        bai += struct.pack("<i", len(bins[t]))
        for b, chunks in sorted(bins[t].items()):
            bai += struct.pack("<Ii", b, len(chunks))
            for vb, ve in chunks:
                bai += struct.pack("<QQ", vb, ve)
        n_intv = (max(lin[t]) + 1) if lin[t] else 0
        bai += struct.pack("<i", n_intv)
        last = 0
        for w in range(n_intv):
            if w in lin[t]:
                last = lin[t][w]
            bai += struct.pack("<Q", last)
    with open(path + ".bai", "wb") as fh:
        fh.write(bai)
    return path


def filter_bam(src: str, dst: str, keep: Callable[[Rec], bool]):
    text, refs, recs = read_bam(src)
    return write_bam(dst, text, refs, [r for r in recs if keep(r)])


def sam_line_to_rec(line: str, refs) -> Rec:
    """SAM text -> BAM record (asBam); supports A, i, Z tags."""
    f = line.rstrip("\n").split("\t")
    names = [n for n, _ in refs]
    qname, flag, rname, pos, mapq, cigar, rnext, pnext, tlen, seq, qual = f[:11]
    tid = names.index(rname) if rname != "*" else -1
    mtid = tid if rnext == "=" else (names.index(rnext) if rnext != "*" else -1)
    ops = []
    num = ""
    for ch in cigar:
        if ch.isdigit():
            num += ch
        elif ch != "*":
            ops.append((int(num) << 4) | _CIGAR_OPS.index(ch))
            num = ""
    l_seq = 0 if seq == "*" else len(seq)
    sb = bytearray((l_seq + 1) // 2)
    for i in range(l_seq):
        sb[i >> 1] |= _NT16_CODE.get(seq[i].upper(), 15) << (4 if i % 2 == 0 else 0)
    qb = bytes([0xFF] * l_seq) if qual == "*" else bytes(ord(c) - 33 for c in qual)
    aux = bytearray()
    for t in f[11:]:
        tag, ty, val = t.split(":", 2)
        if ty == "Z":
            aux += tag.encode() + b"Z" + val.encode() + b"\0"
        elif ty == "A":
            aux += tag.encode() + b"A" + val.encode()
        elif ty == "i":
            v = int(val)
            if 0 <= v < 256:
                aux += tag.encode() + b"C" + struct.pack("<B", v)
            else:
                aux += tag.encode() + b"i" + struct.pack("<i", v)
        else:
            raise ValueError(ty)
    qn = qname.encode() + b"\0"
    p0 = int(pos) - 1
    rlen = sum(v >> 4 for v in ops if (v & 0xF) in (0, 2, 3, 7, 8)) or 1
    core = struct.pack("<iiBBHHHIiii", tid, p0, len(qn), int(mapq), _reg2bin(p0, p0 + rlen), len(ops), int(flag),
                       l_seq, mtid, int(pnext) - 1, int(tlen))
    raw = core + qn + struct.pack(f"<{len(ops)}I", *ops) + bytes(sb) + qb + bytes(aux)
    return Rec(raw)


def sam_to_bam(path: str, header_lines: List[str], aln_lines: List[str]):
    refs = []
    for h in header_lines:
        if h.startswith("@SQ"):
            d = dict(x.split(":", 1) for x in h.split("\t")[1:])
            refs.append((d["SN"], int(d["LN"])))
    recs = [sam_line_to_rec(a, refs) for a in aln_lines]
    return write_bam(path, "\n".join(header_lines) + "\n", refs, recs)
