"""Seeded synthetic RNA-seq-like alignments of the shapes BASELINE.json names.

One generator, written with torch ops so that it runs
  * on the CPU -> numpy arrays -> BAM/BAI/FASTA files (tests, e2e through the .Call layer, CPU baseline), and
  * on the GPU -> HBM-resident read batches in the raer_read_batch layout (bench "value").
Not part of the pileup path. torch is plumbing here (random numbers, sort, gather).

Shape (bulk, config C2): 2x100 bp proper pairs, fragment length N(180, 40) clipped to [100, 400],
strand by 50 kb "gene" blocks consistent with fr-first-strand, CIGAR mix 80% 100M / 12% aMbNcM /
4% soft clip / 2% insertion / 2% deletion, base quality 90% Q37 / 8% Q11-25 / 2% Q2, substitution
errors at 10^(-Q/10), A->G edits at 1e-3 of A positions with level U[0.05, 0.5], MAPQ 255/3/0.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

READ_LEN = 100
_NT16 = {"A": 1, "C": 2, "G": 4, "T": 8, "N": 15}
_CODE_OF_IDX = torch.tensor([1, 2, 4, 8], dtype=torch.uint8)  # A C G T


def make_reference(contig_lens: List[int], seed: int, device="cpu") -> List[torch.Tensor]:
    """Uniform ACGT reference, one uint8 tensor (values 0..3) per contig."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return [torch.randint(0, 4, (L,), generator=g, device=device, dtype=torch.uint8) for L in contig_lens]


def reference_ascii(ref_idx: torch.Tensor) -> torch.Tensor:
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=ref_idx.device)
    return lut[ref_idx.long()]


def write_fasta(path: str, names: List[str], refs: List[torch.Tensor], width: int = 60) -> None:
    with open(path, "w") as fa, open(path + ".fai", "w") as fai:
        off = 0
        for nm, r in zip(names, refs):
            hdr = f">{nm}\n"
            fa.write(hdr)
            off += len(hdr)
            s = reference_ascii(r).cpu().numpy().tobytes().decode()
            L = len(s)
            fai.write(f"{nm}\t{L}\t{off}\t{width}\t{width + 1}\n")
            for i in range(0, L, width):
                fa.write(s[i:i + width] + "\n")
            off += L + (L + width - 1) // width


@dataclass
class Reads:
    """Coordinate-sorted records of one contig of one BAM (device tensors)."""

    pos: torch.Tensor        # int32
    end: torch.Tensor        # int32
    flag: torch.Tensor       # int32
    mapq: torch.Tensor       # uint8
    cigar: torch.Tensor      # [n, 3] uint32-in-int64 (len<<4|op), 0 = unused slot
    n_cigar: torch.Tensor    # int32
    seq: torch.Tensor        # [n, READ_LEN] uint8 4-bit codes
    qual: torch.Tensor       # [n, READ_LEN] uint8
    pair_id: torch.Tensor    # int64
    mate_pos: torch.Tensor   # int32
    isize: torch.Tensor      # int32
    mate_idx: torch.Tensor   # int64 index of the mate inside this contig's sorted order

    def window(self, beg: int, end: int) -> "Reads":
        """The records a tile [beg, end) of a partitioned contig needs: the reads that overlap it (the tile's reads plus
        their halo) and, so that the tiles' rows add up to exactly the rows of the whole contig, the mates of those reads:
        htslib's overlap pass walks both mates from their starts, and its deletion / ref-skip catch-up can change
        qualities of the later mate beyond the earlier mate's end, i.e. inside a tile the earlier mate does not reach.
        (A call with region = "contig:beg-end" through the BAM index -- the reference's own region mode -- does not see
        such a mate either; there the rows of a region can differ from the whole-contig rows at those rare sites.)"""
        keep = (self.end > beg) & (self.pos < end)
        has_mate = self.mate_idx >= 0
        keep = keep | (has_mate & keep[self.mate_idx.clamp(min=0)])
        idx = keep.nonzero().flatten()
        new_of_old = torch.full((self.pos.numel(),), -1, dtype=torch.int64, device=self.pos.device)
        new_of_old[idx] = torch.arange(idx.numel(), device=self.pos.device)
        m = self.mate_idx[idx]
        mate = torch.where(m >= 0, new_of_old[m.clamp(min=0)], torch.full_like(m, -1))
        return Reads(pos=self.pos[idx], end=self.end[idx], flag=self.flag[idx], mapq=self.mapq[idx], cigar=self.cigar[idx],
                     n_cigar=self.n_cigar[idx], seq=self.seq[idx], qual=self.qual[idx], pair_id=self.pair_id[idx],
                     mate_pos=self.mate_pos[idx], isize=self.isize[idx], mate_idx=mate)


def gen_contig(ref: torch.Tensor, n_pairs: int, seed: int, *, edit_frac=1e-3, gene_block=50_000,
               p_splice=0.12, p_clip=0.04, p_ins=0.02, p_del=0.02, intron=(100, 5000), device="cpu",
               qual_mix=(0.90, 0.08, 0.02), mapq_mix=(0.95, 0.03, 0.02), single_end=False,
               starts: Optional[torch.Tensor] = None, read_len: int = READ_LEN) -> Reads:
    """2x100 bp pairs on one contig (or single-end 91 bp style reads when single_end). starts: fragment starts to use
    instead of uniform ones (n_pairs int64 values on `device`; the C4 workload puts 70 % of them inside intervals)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = ref.numel()
    n = n_pairs
    U = lambda *s: torch.rand(*s, generator=g, device=dev)  # noqa: E731
    frag = (180 + 40 * torch.randn(n, generator=g, device=dev)).round().clamp(100, 400).long()
    if single_end:
        frag = torch.full((n,), read_len, device=dev, dtype=torch.long)
    max_span = read_len + intron[1] + 8
    start = (U(n) * max(L - 400 - max_span, 1)).long()
    if starts is not None:
        start = starts.to(dev).long().clamp(0, max(L - 400 - max_span, 1) - 1)
    # left mate starts at the fragment start, right mate ends at the fragment end
    lpos = start
    rpos = start + frag - read_len
    gene_minus = ((start // gene_block) % 2 == 1)
    n_reads = n if single_end else 2 * n
    pos0 = lpos if single_end else torch.cat([lpos, rpos])
    is_right = torch.zeros(n_reads, dtype=torch.bool, device=dev)
    if not single_end:
        is_right[n:] = True
    pid = torch.arange(n, device=dev).repeat(1 if single_end else 2)
    gm = gene_minus.repeat(1 if single_end else 2)
    # fr-first-strand: read1 is antisense. plus gene: left=R2 fwd (163), right=R1 rev (83); minus gene: left=R1 fwd (99), right=R2 rev (147)
    if single_end:
        flag = torch.where(gm, torch.tensor(16, device=dev), torch.tensor(0, device=dev))
    else:
        flag = torch.where(is_right, torch.where(gm, torch.tensor(147, device=dev), torch.tensor(83, device=dev)),
                           torch.where(gm, torch.tensor(99, device=dev), torch.tensor(163, device=dev)))
    # CIGAR type per read
    u = U(n_reads)
    c1, c2, c3, c4 = p_splice, p_splice + p_clip, p_splice + p_clip + p_ins, p_splice + p_clip + p_ins + p_del
    ctype = torch.zeros(n_reads, dtype=torch.long, device=dev)
    ctype[u < c4] = 4
    ctype[u < c3] = 3
    ctype[u < c2] = 2
    ctype[u < c1] = 1
    a = (10 + U(n_reads) * 80).long()                       # split point
    k = (1 + U(n_reads) * 3).long().clamp(1, 3)             # indel / clip length base
    nlen = (intron[0] + U(n_reads) * (intron[1] - intron[0])).long()
    clip_left = U(n_reads) < 0.5
    k_clip = (1 + U(n_reads) * 20).long()
    j = torch.arange(read_len, device=dev).unsqueeze(0)     # [1, RL]
    A = a.unsqueeze(1)
    K = k.unsqueeze(1)
    KC = k_clip.unsqueeze(1)
    NL = nlen.unsqueeze(1)
    P0 = pos0.unsqueeze(1)
    T = ctype.unsqueeze(1)
    CL = clip_left.unsqueeze(1)
    # reference offset of query base j (or -1 when the base is clipped / inserted)
    off = j.expand(n_reads, read_len).clone()
    off = torch.where(T == 1, j + (j >= A) * NL, off)
    off = torch.where((T == 2) & CL, torch.where(j < KC, torch.full_like(off, -1), j - KC), off)
    off = torch.where((T == 2) & ~CL, torch.where(j >= read_len - KC, torch.full_like(off, -1), off), off)
    off = torch.where(T == 3, torch.where(j < A, j, torch.where(j < A + K, torch.full_like(off, -1), j - K)), off)
    off = torch.where(T == 4, j + (j >= A) * K, off)
    refpos = torch.where(off >= 0, P0 + off, torch.full_like(off, -1))
    refpos = refpos.clamp(max=L - 1)
    span = torch.full((n_reads,), read_len, device=dev, dtype=torch.long)
    span = torch.where(ctype == 1, read_len + nlen, span)
    span = torch.where(ctype == 2, read_len - k_clip, span)
    span = torch.where(ctype == 3, read_len - k, span)
    span = torch.where(ctype == 4, read_len + k, span)
    # CIGAR words (len << 4 | op): M0 I1 D2 N3 S4
    w = lambda ln, op: (ln << 4) | op  # noqa: E731
    z = torch.zeros(n_reads, dtype=torch.long, device=dev)
    cig = torch.stack([w(torch.full_like(z, read_len), 0), z, z], dim=1)
    m = ctype == 1
    cig[m] = torch.stack([w(a, 0), w(nlen, 3), w(read_len - a, 0)], dim=1)[m]
    m = (ctype == 2) & clip_left
    cig[m] = torch.stack([w(k_clip, 4), w(read_len - k_clip, 0), z], dim=1)[m]
    m = (ctype == 2) & ~clip_left
    cig[m] = torch.stack([w(read_len - k_clip, 0), w(k_clip, 4), z], dim=1)[m]
    m = ctype == 3
    cig[m] = torch.stack([w(a, 0), w(k, 1), w(read_len - a - k, 0)], dim=1)[m]
    m = ctype == 4
    cig[m] = torch.stack([w(a, 0), w(k, 2), w(read_len - a, 0)], dim=1)[m]
    n_cig = torch.where(ctype == 0, 1, torch.where(ctype == 2, 2, 3)).int()
    # bases: reference, then A->G editing (T->C for minus genes), then sequencing errors
    base = ref[refpos.clamp(min=0)].long()
    rnd_base = (U(n_reads, read_len) * 4).long().clamp(0, 3)
    base = torch.where(refpos >= 0, base, rnd_base)
    g2 = torch.Generator(device=dev)
    g2.manual_seed(seed * 7919 + 13)
    site_u = torch.rand(L, generator=g2, device=dev)
    level = 0.05 + 0.45 * torch.rand(L, generator=g2, device=dev)
    editable = site_u < edit_frac
    rp = refpos.clamp(min=0)
    ed_hit = editable[rp] & (refpos >= 0) & (U(n_reads, read_len) < level[rp])
    GM = gm.unsqueeze(1)
    base = torch.where(ed_hit & ~GM & (base == 0), torch.full_like(base, 2), base)   # A->G on plus genes
    base = torch.where(ed_hit & GM & (base == 3), torch.full_like(base, 1), base)    # T->C on minus genes
    uq = U(n_reads, read_len)
    qual = torch.full((n_reads, read_len), 37, dtype=torch.long, device=dev)
    qual = torch.where(uq < qual_mix[1] + qual_mix[2], (11 + U(n_reads, read_len) * 15).long(), qual)
    qual = torch.where(uq < qual_mix[2], torch.full_like(qual, 2), qual)
    perr = torch.pow(10.0, -qual.float() / 10.0)
    err = U(n_reads, read_len) < perr
    base = torch.where(err, (base + 1 + (U(n_reads, read_len) * 3).long().clamp(0, 2)) % 4, base)
    codes = _CODE_OF_IDX.to(dev)[base]
    um = U(n)
    mq = torch.full((n,), 255, dtype=torch.long, device=dev)
    mq = torch.where(um < mapq_mix[1] + mapq_mix[2], torch.full_like(mq, 3), mq)
    mq = torch.where(um < mapq_mix[2], torch.zeros_like(mq), mq)
    mapq = mq.repeat(1 if single_end else 2)
    end = pos0 + span
    # mates
    if single_end:
        mate_pos = torch.full((n_reads,), -1, device=dev, dtype=torch.long)
        isize = torch.zeros(n_reads, device=dev, dtype=torch.long)
    else:
        mate_pos = torch.cat([rpos, lpos])
        tl = (end[n:] - lpos)
        isize = torch.cat([tl, -tl])
    order = torch.argsort(pos0 * 2 + is_right.long(), stable=True)
    inv = torch.empty_like(order)
    inv[order] = torch.arange(n_reads, device=dev)
    if single_end:
        mate_idx = torch.full((n_reads,), -1, device=dev, dtype=torch.long)
    else:
        partner = torch.cat([torch.arange(n, 2 * n, device=dev), torch.arange(0, n, device=dev)])
        mate_idx = inv[partner][order]
    return Reads(
        pos=pos0[order].int(), end=end[order].int(), flag=flag[order].int(), mapq=mapq[order].to(torch.uint8),
        cigar=cig[order], n_cigar=n_cig[order], seq=codes[order], qual=qual[order].to(torch.uint8),
        pair_id=pid[order], mate_pos=mate_pos[order].int(), isize=isize[order].int(), mate_idx=mate_idx,
    )


# --------------------------------------------------------------------------------------- BAM files
_SYNTH = None


def _synth_lib():
    global _SYNTH
    if _SYNTH is None:
        so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libraer_synth.so")
        L = C.CDLL(so)
        L.synth_write_bam.restype = C.c_int
        _SYNTH = L
    return _SYNTH


def write_bam(path: str, names: List[str], lens: List[int], per_contig: List[Optional[Reads]], *, qname_prefix="r",
              tags: Optional[Dict[str, np.ndarray]] = None, level: int = 1) -> int:
    """Write one BAM (+.bai) from per-contig Reads. Returns the number of records."""
    tid, pos, flag, mapq, mpos, isize, ncig, cig, seq, qual, pid = [], [], [], [], [], [], [], [], [], [], []
    for t, r in enumerate(per_contig):
        if r is None or r.pos.numel() == 0:
            continue
        k = r.pos.numel()
        tid.append(np.full(k, t, dtype=np.int32))
        pos.append(r.pos.cpu().numpy()); flag.append(r.flag.cpu().numpy().astype(np.uint16))
        mapq.append(r.mapq.cpu().numpy()); mpos.append(r.mate_pos.cpu().numpy()); isize.append(r.isize.cpu().numpy())
        ncig.append(r.n_cigar.cpu().numpy()); cig.append(r.cigar.cpu().numpy()); seq.append(r.seq.cpu().numpy())
        qual.append(r.qual.cpu().numpy()); pid.append(r.pair_id.cpu().numpy() + (t << 40))
    if not tid:
        tid = [np.zeros(0, np.int32)]
    cat = np.concatenate
    tid, pos, flag, mapq, mpos, isize, ncig = map(cat, (tid, pos, flag, mapq, mpos, isize, ncig))
    cig, seq, qual, pid = cat(cig), cat(seq), cat(qual), cat(pid)
    n = len(tid)
    cig_flat = cig[np.arange(3)[None, :] < ncig[:, None]].astype(np.uint32)
    cig_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(ncig, out=cig_off[1:])
    rl = seq.shape[1] if n else READ_LEN
    sq_off = np.arange(n + 1, dtype=np.int64) * rl
    qn = [f"{qname_prefix}{int(p)}".encode() for p in pid]
    qn_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(q) for q in qn], out=qn_off[1:])
    qn_blob = b"".join(qn)
    if tags:
        blobs = []
        for i in range(n):
            b = b""
            for tg, vals in tags.items():
                v = vals[i]
                if v is not None and v != "":
                    b += tg.encode() + b"Z" + str(v).encode() + b"\0"
            blobs.append(b)
        aux_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum([len(b) for b in blobs], out=aux_off[1:])
        aux_blob = b"".join(blobs)
    else:
        aux_off, aux_blob = None, b""
    mtid = np.where(mpos >= 0, tid, -1).astype(np.int32)
    hdr = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{nm}\tLN:{ln}\n" for nm, ln in zip(names, lens))
    names_c = (C.c_char_p * len(names))(*[x.encode() for x in names])
    lens_c = (C.c_int32 * len(lens))(*lens)
    P = lambda a, t: np.ascontiguousarray(a).ctypes.data_as(C.POINTER(t))  # noqa: E731
    seq_c = np.ascontiguousarray(seq.reshape(-1))
    qual_c = np.ascontiguousarray(qual.reshape(-1))
    rc = _synth_lib().synth_write_bam(
        path.encode(), hdr.encode(), len(names), names_c, lens_c, C.c_int64(n), P(tid, C.c_int32), P(pos, C.c_int32),
        P(flag, C.c_uint16), P(mapq, C.c_uint8), P(mtid, C.c_int32), P(mpos.astype(np.int32), C.c_int32),
        P(isize.astype(np.int32), C.c_int32), P(cig_off, C.c_int64), P(cig_flat, C.c_uint32), P(sq_off, C.c_int64),
        P(seq_c, C.c_uint8), P(qual_c, C.c_uint8), P(qn_off, C.c_int64), C.c_char_p(qn_blob),
        P(aux_off, C.c_int64) if aux_off is not None else None, C.c_char_p(aux_blob) if aux_off is not None else None,
        C.c_int(level))
    if rc != 0:
        raise RuntimeError("synth_write_bam failed")
    return n


def make_bulk_dataset(outdir: str, *, n_bams=2, n_contigs=2, contig_len=200_000, pairs_per_contig=20_000, seed=1001,
                      **gen_kw) -> dict:
    """Reference + n_bams BAMs of the C2 shape at a chosen size. Returns paths and counts."""
    os.makedirs(outdir, exist_ok=True)
    names = [f"chr{i + 1}" for i in range(n_contigs)]
    lens = [contig_len] * n_contigs
    refs = make_reference(lens, seed)
    fa = os.path.join(outdir, "ref.fa")
    write_fasta(fa, names, refs)
    from concurrent.futures import ThreadPoolExecutor

    def one(b):  # the deflate loop of the writer runs outside the GIL: the BAMs are written side by side
        per = [gen_contig(refs[t], pairs_per_contig, seed + 1000 * (b + 1) + t, **gen_kw) for t in range(n_contigs)]
        p = os.path.join(outdir, f"s{b + 1}.bam")
        write_bam(p, names, lens, per, qname_prefix=f"b{b}_")
        nb = 0
        for r in per:
            c = r.cigar
            op, ln = c & 0xF, c >> 4
            nb += int((ln * (op == 0)).sum())
        return p, nb

    with ThreadPoolExecutor(max_workers=max(1, min(n_bams, 4))) as ex:
        done = list(ex.map(one, range(n_bams)))
    bams = [p for p, _ in done]
    n_bases = sum(nb for _, nb in done)
    return {"fasta": fa, "bams": bams, "names": names, "lens": lens, "aligned_bases": n_bases}


def known_sites_mask(ref: torch.Tensor, seed: int, frac: float = 0.4, gene_block: int = 50_000) -> torch.Tensor:
    """C3 site list as a bool tensor over the contig: A positions of plus genes / T positions of minus genes (the
    positions A-to-I editing can show at, seen from the plus strand), a fraction `frac` of them (0.4 -> 10 % of all
    positions: 10^7 sites on the 100 Mb reference)."""
    g = torch.Generator(device=ref.device)
    g.manual_seed(seed)
    L = ref.numel()
    minus = (torch.arange(L, device=ref.device) // gene_block) % 2 == 1
    cand = torch.where(minus, ref == 3, ref == 0)
    return cand & (torch.rand(L, generator=g, device=ref.device) < frac)


def alu_intervals(L: int, seed: int, device="cpu", slot: int = 1000, len_lo: int = 280, len_hi: int = 320):
    """C4 intervals of one contig: one per `slot` positions, non-overlapping, length U[len_lo, len_hi]
    (10^6 of them on the 1 Gb reference). Returns (start0, end0_exclusive) int64 tensors."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = L // slot
    ln = torch.randint(len_lo, len_hi + 1, (n,), generator=g, device=device)
    st = torch.arange(n, device=device) * slot + (torch.rand(n, generator=g, device=device) * (slot - len_hi)).long()
    return st, st + ln


def interval_mask(L: int, st: torch.Tensor, en: torch.Tensor) -> torch.Tensor:
    d = torch.zeros(L + 1, dtype=torch.int32, device=st.device)
    d.index_add_(0, st, torch.ones_like(st, dtype=torch.int32))
    d.index_add_(0, en, -torch.ones_like(en, dtype=torch.int32))
    return torch.cumsum(d[:L], 0) > 0


def interval_starts(n_pairs: int, L: int, st: torch.Tensor, en: torch.Tensor, seed: int, inside: float = 0.7) -> torch.Tensor:
    """fragment starts of the C4 BAM: `inside` of the fragments start within 60 bases in front of an interval or
    inside it, the rest anywhere"""
    dev = st.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    pick = torch.randint(0, st.numel(), (n_pairs,), generator=g, device=dev)
    ln = (en - st)[pick]
    s_in = st[pick] - 60 + (torch.rand(n_pairs, generator=g, device=dev) * (ln + 60).float()).long()
    s_any = (torch.rand(n_pairs, generator=g, device=dev) * max(L - 6000, 1)).long()
    return torch.where(torch.rand(n_pairs, generator=g, device=dev) < inside, s_in.clamp(min=0), s_any)


def pack_bitmap(mask: torch.Tensor) -> torch.Tensor:
    """bool tensor -> uint32 words (as int32), bit (p & 31) of word p >> 5: the site_mask layout of raer_bulk_conf"""
    L = mask.numel()
    pad = (-L) % 32
    if pad:
        mask = torch.cat([mask, torch.zeros(pad, dtype=torch.bool, device=mask.device)])
    w = (mask.view(-1, 32).long() << torch.arange(32, device=mask.device)).sum(1)
    return (w & 0xFFFFFFFF).to(torch.int64).where(w < 2**31, w - 2**32).to(torch.int32).contiguous()


def make_aei_dataset(outdir: str, *, n_contigs=2, contig_len=200_000, pairs_per_contig=20_000, seed=4001) -> dict:
    """Reference, one BAM of the C4 shape (70 % of the fragments at the intervals) and the interval list."""
    os.makedirs(outdir, exist_ok=True)
    names = [f"chr{i + 1}" for i in range(n_contigs)]
    lens = [contig_len] * n_contigs
    refs = make_reference(lens, seed)
    fa = os.path.join(outdir, "ref.fa")
    write_fasta(fa, names, refs)
    per, n_bases = [], 0
    sn, ss, se = [], [], []
    for t in range(n_contigs):
        st, en = alu_intervals(contig_len, seed + 10 + t)
        starts = interval_starts(pairs_per_contig, contig_len, st, en, seed + 100 + t)
        r = gen_contig(refs[t], pairs_per_contig, seed + 1000 + t, starts=starts)
        per.append(r)
        c = r.cigar
        n_bases += int(((c >> 4) * ((c & 0xF) == 0)).sum())
        sn += [names[t]] * st.numel()
        ss += (st + 1).tolist()
        se += en.tolist()
    bam = os.path.join(outdir, "aei.bam")
    write_bam(bam, names, lens, per, qname_prefix="a_")
    return {"fasta": fa, "bam": bam, "names": names, "lens": lens, "aligned_bases": n_bases,
            "intervals": dict(seqnames=sn, start=ss, end=se)}


def make_sc_dataset(outdir: str, *, n_contigs=2, contig_len=60_000, reads_per_contig=20_000, n_cb=60, n_sites=400,
                    seed=5001, gene_block=5_000, edit_frac=0.02) -> dict:
    """Reference, one 10x-shaped BAM and a site list of the C5 shape at a chosen size: single-end
    reads, flags 0/16 by gene block (+1024 on 20 %), CB:Z drawn Zipf(s=1) from `n_cb` whitelist barcodes with 1 %
    of the reads without the tag and 1 % off the list, UB:Z 12-mers with about 3 reads per (CB, UMI, 150-base locus)
    (2 % without the tag), sites = A positions of plus genes / T positions of minus genes (REF A, ALT G in the
    strand's own orientation) plus a few C>T sites."""
    os.makedirs(outdir, exist_ok=True)
    rng = np.random.default_rng(seed)
    names = [f"chr{i + 1}" for i in range(n_contigs)]
    lens = [contig_len] * n_contigs
    refs = make_reference(lens, seed)
    fa = os.path.join(outdir, "ref.fa")
    write_fasta(fa, names, refs)
    acgt = np.array(list("ACGT"))
    whitelist = ["".join(acgt[rng.integers(0, 4, 16)]) + "-1" for _ in range(n_cb)]
    zipf = 1.0 / np.arange(1, n_cb + 1)
    zipf /= zipf.sum()
    per, cbs, ubs = [], [], []
    for t in range(n_contigs):
        r = gen_contig(refs[t], reads_per_contig, seed + 17 * (t + 1), single_end=True, gene_block=gene_block,
                       edit_frac=edit_frac)
        n = r.pos.numel()
        dup = torch.from_numpy(rng.random(n) < 0.2)
        r.flag = r.flag | (dup.int() * 1024)
        per.append(r)
        cb_idx = rng.choice(n_cb, size=n, p=zipf)
        # reads of one molecule start within 150 bases of each other, so that their sites overlap
        locus = r.pos.numpy() // 150
        key = cb_idx.astype(np.int64) * 1_000_000 + locus
        _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
        k = np.maximum(cnt[inv] // 3, 1)
        u = (rng.random(n) * k).astype(np.int64)
        h = (key * 1_000_003 + u * 7919 + 12345) % (1 << 24)
        u_cb, u_ub = rng.random(n), rng.random(n)
        for i in range(n):
            if u_cb[i] < 0.01:
                cbs.append(None)
            elif u_cb[i] < 0.02:
                cbs.append("".join(acgt[rng.integers(0, 4, 16)]) + "-9")
            else:
                cbs.append(whitelist[cb_idx[i]])
            ubs.append(None if u_ub[i] < 0.02 else "".join(acgt[(int(h[i]) >> (2 * j)) & 3] for j in range(12)))
    bam = os.path.join(outdir, "sc.bam")
    write_bam(bam, names, lens, per, qname_prefix="sc_", tags={"CB": cbs, "UB": ubs})
    sn, sp, ss, sr, sa = [], [], [], [], []
    for t in range(n_contigs):
        ref = refs[t].numpy()
        # half of the candidates are positions gen_contig edits (same generator and seed as its `editable` mask)
        g2 = torch.Generator()
        g2.manual_seed((seed + 17 * (t + 1)) * 7919 + 13)
        ed = np.flatnonzero((torch.rand(contig_len, generator=g2) < edit_frac).numpy())
        ed = ed[(ed >= 200) & (ed < contig_len - 200)]
        rnd = rng.choice(np.arange(200, contig_len - 200), size=min(4 * n_sites, contig_len - 400), replace=False)
        pos = np.unique(np.concatenate([ed, rnd]))
        took = 0
        for p0 in pos:
            minus = (p0 // gene_block) % 2 == 1
            b = ref[p0]
            if (not minus and b == 0) or (minus and b == 3):
                ra = ("A", "G")
            elif took % 7 == 0 and ((not minus and b == 1) or (minus and b == 2)):
                ra = ("C", "T")
            else:
                continue
            sn.append(names[t]); sp.append(int(p0) + 1); ss.append("-" if minus else "+"); sr.append(ra[0]); sa.append(ra[1])
            took += 1
            if took >= n_sites:
                break
    return {"fasta": fa, "bam": bam, "barcodes": whitelist, "names": names,
            "sites": dict(seqnames=sn, start=sp, strand=ss, ref=sr, alt=sa)}


# --------------------------------------------------------------------------------------- device batches
def qname_coin(qname: bytes) -> bool:
    """htslib >= 1.13 overlap tie-break: Wang_hash(X31_hash(qname)) & 1 (true: the earlier mate keeps its qualities)."""
    h = qname[0]
    if h > 127:
        h -= 256
    h &= 0xFFFFFFFF
    for c in qname[1:]:
        h = ((h << 5) - h + (c if c < 128 else c - 256)) & 0xFFFFFFFF
    k = h
    k = (k + (~(k << 15) & 0xFFFFFFFF)) & 0xFFFFFFFF
    k ^= k >> 10
    k = (k + (k << 3)) & 0xFFFFFFFF
    k ^= k >> 6
    k = (k + (~(k << 11) & 0xFFFFFFFF)) & 0xFFFFFFFF
    k ^= k >> 16
    return bool(k & 1)


def pack_device_batch(per_file: List[Reads], tid: int, beg: int, end: int, remove_overlaps: bool = True,
                      qname_prefixes: Optional[List[str]] = None, libtype: int = 1, cb: Optional[torch.Tensor] = None,
                      umi: Optional[torch.Tensor] = None):
    """Lay out the reads of one contig (all files) as a raer_read_batch in the memory of their
    device (HBM when generated on cuda). Returns (dict of tensors kept alive, fields for ReadBatch).
    qname_prefixes: per file, the prefix write_bam() gives the read names of these records; the mate-overlap coin is
    then htslib's hash of the real name "<prefix><pair_id + (tid << 40)>" instead of the integer stand-in (tests
    that hold a packed batch against the oracle on the BAM written from the same records)."""
    from . import _cabi

    dev = per_file[0].pos.device
    hdrs, cigs, seqs, quals, maxends, pairs = [], [], [], [], [], []
    file_off = [0]
    ro = co = so = 0
    for r in per_file:
        n = r.pos.numel()
        rl = r.seq.shape[1]
        rl_pad = (rl + 7) & ~7  # reads start on 4-byte sequence / 8-byte quality boundaries (raer_read_hdr.sq_off)
        h = torch.zeros((n, 8), dtype=torch.int32, device=dev)
        h[:, 0] = r.pos
        h[:, 1] = r.end
        bits = torch.zeros(n, dtype=torch.int32, device=dev)
        # fr-first-strand inversion (src/plp_utils.c:339-352): read1 forward or read2 reverse (unpaired: forward)
        fl = r.flag
        rev = (fl & 16) != 0
        paired = (fl & 1) != 0
        r1 = (~paired) | ((fl & 64) != 0)
        inv = torch.where(r1, ~rev, rev)  # fr-first-strand
        if libtype == 2:
            inv = ~inv
        elif libtype == 0:
            inv = torch.zeros_like(inv)
        bits |= inv.int() * 1
        h[:, 2] = (fl & 0xFFFF) | (r.mapq.int() << 16) | (bits << 24)
        h[:, 3] = rl | (r.n_cigar << 16)
        ncum = torch.cumsum(r.n_cigar.long(), 0)
        h[:, 4] = (co + ncum - r.n_cigar.long()).int()
        h[:, 5] = (so + torch.arange(n, device=dev) * (rl_pad // 2)).int()
        mate = torch.full((n,), -1, dtype=torch.int64, device=dev)
        if remove_overlaps and (r.mate_idx >= 0).any():
            idx = torch.arange(n, device=dev)
            m = r.mate_idx
            later = (m >= 0) & (m < idx)                     # the mate arrived first
            mpos = r.pos[m.clamp(min=0)]
            mend = r.end[m.clamp(min=0)]
            # overlap_push: skip pairs that cannot overlap (|isize| >= 2*l_qseq and mate start beyond own end)
            ok = later & ~((r.isize.abs() >= 2 * rl) & (r.mate_pos >= r.end))
            ok &= mend >= r.pos                              # entry still alive in the pileup buffer
            mate = torch.where(ok, m + ro, mate)
            if qname_prefixes is not None:
                pfx = qname_prefixes[len(hdrs)]
                pids = (r.pair_id[m.clamp(min=0)] + (tid << 40)).cpu().numpy()
                okn = ok.cpu().numpy()
                coin = np.zeros(n, dtype=bool)
                for i in np.flatnonzero(okn):
                    coin[i] = qname_coin(f"{pfx}{int(pids[i])}".encode())
                keep_a = torch.from_numpy(coin).to(dev)
            else:
                keep_a = _keep_a_bits(r.pair_id[m.clamp(min=0)])
            h[:, 2] |= (ok & keep_a).int() << (24 + 3)
            pairs.append((idx[ok] + ro).int())
            del mpos
        h[:, 6] = mate.int()
        h[:, 7] = (torch.arange(n, device=dev) + ro).int()
        hdrs.append(h)
        valid = torch.arange(3, device=dev).unsqueeze(0) < r.n_cigar.unsqueeze(1)
        cigs.append(r.cigar[valid].int())
        sq = r.seq
        if rl_pad != rl:
            sq = torch.cat([sq, torch.full((n, rl_pad - rl), 15, dtype=torch.uint8, device=dev)], 1)
            ql = torch.cat([r.qual, torch.zeros((n, rl_pad - rl), dtype=torch.uint8, device=dev)], 1)
        else:
            ql = r.qual
        seqs.append(((sq[:, 0::2] << 4) | sq[:, 1::2]).reshape(-1))
        quals.append(ql.reshape(-1))
        maxends.append(torch.cummax(r.end, 0)[0].int())
        ro += n
        co += int(ncum[-1]) if n else 0
        so += n * (rl_pad // 2)
        file_off.append(ro)
    T = {
        "hdr": torch.cat(hdrs).contiguous(), "cigar": torch.cat(cigs).contiguous(), "seq": torch.cat(seqs).contiguous(),
        "qual": torch.cat(quals).contiguous(), "maxend": torch.cat(maxends).contiguous(),
        "pair_b": (torch.cat(pairs) if pairs else torch.zeros(0, dtype=torch.int32, device=dev)).contiguous(),
        "file_off": np.asarray(file_off, dtype=np.int64),
    }
    b = _cabi.ReadBatch()
    b.n_files = len(per_file)
    b.tid, b.beg, b.end = tid, beg, end
    b.n_reads = ro
    b.file_off = T["file_off"].ctypes.data
    b.hdr, b.maxend, b.cigar = T["hdr"].data_ptr(), T["maxend"].data_ptr(), T["cigar"].data_ptr()
    b.seq, b.qual = T["seq"].data_ptr(), T["qual"].data_ptr()
    b.n_cigar_words, b.n_sq_bytes = co, so
    b.pair_b, b.n_pairs = T["pair_b"].data_ptr(), T["pair_b"].numel()
    if cb is not None:  # single cell: whitelist column (1-based, 0 = none) and UMI id (0 = no tag) per read
        T["cb"] = cb.to(torch.int32).contiguous()
        T["umi"] = (umi if umi is not None else torch.zeros_like(cb)).to(torch.int32).contiguous()
        b.cb, b.umi = T["cb"].data_ptr(), T["umi"].data_ptr()
    return T, b


def gen_sc_batch(ref: torch.Tensor, n_reads: int, seed: int, *, n_cb=10_000, n_site_frac=0.01, read_len=91, gene_block=50_000,
                 tid=0):
    """One C5-shaped batch on the device of `ref`: single-end reads (flags 0 / 16 by gene block, +1024 on 20 %), CB
    drawn Zipf(s = 1) from n_cb barcodes with 2 % of the reads off the list, UMI ids with about 3 reads per (barcode,
    150-base locus, UMI), and the site list (A positions of plus genes / T positions of minus genes, REF A / ALT G in
    the strand's own orientation, n_site_frac of all positions). Returns (tensors, ReadBatch, sites dict)."""
    dev = ref.device
    L = ref.numel()
    r = gen_contig(ref, n_reads, seed, device=dev, single_end=True, gene_block=gene_block, read_len=read_len, edit_frac=0.02)
    g = torch.Generator(device=dev)
    g.manual_seed(seed * 31 + 5)
    n = r.pos.numel()
    r.flag = r.flag | ((torch.rand(n, generator=g, device=dev) < 0.2).int() * 1024)
    zipf = 1.0 / torch.arange(1, n_cb + 1, device=dev, dtype=torch.float64)
    cb_idx = torch.multinomial((zipf / zipf.sum()).float(), n, replacement=True, generator=g)
    cb = torch.where(torch.rand(n, generator=g, device=dev) < 0.02, torch.zeros_like(cb_idx), cb_idx + 1)
    locus = r.pos.long() // 150
    key = cb_idx * 1_000_000 + locus
    u = (torch.rand(n, generator=g, device=dev) * 3).long()
    umi = (key * 1_000_003 + u * 7919 + 12345) % (1 << 24)
    umi = torch.where(torch.rand(n, generator=g, device=dev) < 0.02, torch.zeros_like(umi), umi + 1)
    T, b = pack_device_batch([r], tid, 0, L, remove_overlaps=False, libtype=2, cb=cb, umi=umi)
    minus = (torch.arange(L, device=dev) // gene_block) % 2 == 1
    cand = torch.where(minus, ref == 3, ref == 0)
    sel = cand & (torch.rand(L, generator=g, device=dev) < n_site_frac * 4)
    sel[:200] = False
    sel[L - 200:] = False
    pos = sel.nonzero().flatten()
    T["site_pos"] = pos.to(torch.int32).contiguous()
    T["site_strand"] = torch.where(minus[pos], 2, 1).to(torch.uint8).contiguous()
    T["site_ref"] = torch.full((pos.numel(),), ord("A"), dtype=torch.uint8, device=dev)
    T["site_alt"] = torch.full((pos.numel(),), ord("G"), dtype=torch.uint8, device=dev)
    c = r.cigar
    T["aligned_bases"] = int(((c >> 4) * ((c & 0xF) == 0)).sum())
    return T, b


def _keep_a_bits(pair_id: torch.Tensor) -> torch.Tensor:
    """Stand-in for htslib's qname-hash coin (Wang(X31(qname)) & 1) on device-generated batches that
    have no qname strings: a fixed integer hash of the pair id. Only used by bench data."""
    x = (pair_id * 2654435761) & 0xFFFFFFFF
    x = x ^ (x >> 15)
    return (x & 1) == 1
