"""The batch layer of the C ABI (include/raer_b200.h, layer 2) on device-generated read batches:
the plain host path, the pipelined host path (raer_pipe_*), both tile kernels, and the two conditions
the fast kernel hands back to the host (tile lists larger than their first allocation, coverage beyond
its 16-bit counters) must all give the same rows, bit for bit."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _conf(n_files, **kw):
    from raer_b200 import _cabi

    conf = _cabi.BulkConf()
    conf.n_files = n_files
    conf.min_depth, conf.min_bq = 1, 20
    conf.report_multiallelic, conf.remove_overlaps, conf.want_stats = 1, 1, 1
    for f in range(n_files):
        conf.min_mapq[f] = 0
        conf.only_keep_variants[f] = kw.get("okv", 1)
    conf.reg_beg, conf.reg_end = 0, 2**31 - 1
    return conf


def _host_batch(n_files, contig_len, pairs, seed):
    """(pinned host tensors, ReadBatch with host pointers) of one synthetic contig"""
    import torch

    from raer_b200 import _cabi, synth

    dev = torch.device("cuda", 0)
    ref = synth.make_reference([contig_len], seed, device=dev)[0]
    per = [synth.gen_contig(ref, pairs, seed + 1000 * (b + 1), device=dev) for b in range(n_files)]
    T, rb = synth.pack_device_batch(per, 0, 0, contig_len)
    T["ref"] = synth.reference_ascii(ref)
    H = {k: torch.empty(T[k].shape, dtype=T[k].dtype, pin_memory=True) for k in
         ("hdr", "maxend", "cigar", "seq", "qual", "pair_b", "ref")}
    for k in H:
        H[k].copy_(T[k])
    torch.cuda.synchronize()
    hb = _cabi.ReadBatch()
    C.memmove(C.byref(hb), C.byref(rb), C.sizeof(hb))
    hb.hdr, hb.maxend, hb.cigar = H["hdr"].data_ptr(), H["maxend"].data_ptr(), H["cigar"].data_ptr()
    hb.seq, hb.qual, hb.pair_b = H["seq"].data_ptr(), H["qual"].data_ptr(), H["pair_b"].data_ptr()
    H["file_off"] = T["file_off"]
    hb.file_off = H["file_off"].ctypes.data
    return H, hb


def _rows_to_np(rows, nf):
    n = rows.n_rows
    if n == 0:
        return dict(pos=np.zeros(0, np.int32), meta=np.zeros(0, np.uint32), stats=np.zeros((0, 3)),
                    counts=np.zeros((0, nf, 8), np.int32), alt=np.zeros((0, nf), np.uint32))
    return dict(pos=np.ctypeslib.as_array(rows.pos, (n,)).copy(), meta=np.ctypeslib.as_array(rows.meta, (n,)).copy(),
                stats=np.ctypeslib.as_array(rows.stats, (n, 3)).copy(),
                counts=np.ctypeslib.as_array(rows.counts, (n, nf, 8)).copy(),
                alt=np.ctypeslib.as_array(rows.altcode, (n, nf)).copy())


def _same(a, b):
    assert a["pos"].shape == b["pos"].shape
    for k in ("pos", "meta", "counts", "alt"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["stats"], b["stats"], equal_nan=True), "stats"


def _run_host(H, hb, conf, nf):
    from raer_b200 import _cabi

    L = _cabi.lib()
    ctx = L.raer_ctx_create(0)
    assert ctx, _cabi.last_error()
    try:
        assert L.raer_ctx_set_reference(ctx, 0, C.c_char_p(H["ref"].data_ptr()), H["ref"].numel()) == 0
        rows = _cabi.Rows()
        rc = L.raer_batch_pileup_host(ctx, C.byref(conf), C.byref(hb), C.byref(rows))
        assert rc == 0, _cabi.last_error()
        out = _rows_to_np(rows, nf)
        L.raer_rows_free(C.byref(rows))
        return out
    finally:
        L.raer_ctx_destroy(ctx)


@pytest.fixture(scope="module")
def small_batch():
    return _host_batch(2, 60_000, 30_000, 77)


def test_fast_and_generic_kernels_agree_on_a_batch(small_batch, monkeypatch):
    H, hb = small_batch
    conf = _conf(2)
    fast = _run_host(H, hb, conf, 2)
    assert fast["pos"].size > 1000
    monkeypatch.setenv("RAER_KERNEL", "generic")
    _same(fast, _run_host(H, hb, conf, 2))
    # without only_keep_variants every covered site is a row
    monkeypatch.delenv("RAER_KERNEL")
    conf2 = _conf(2, okv=0)
    fast2 = _run_host(H, hb, conf2, 2)
    assert fast2["pos"].size > fast["pos"].size
    monkeypatch.setenv("RAER_KERNEL", "generic")
    _same(fast2, _run_host(H, hb, conf2, 2))


def test_pipelined_host_path_equals_plain_host_path(small_batch):
    from raer_b200 import _cabi

    H, hb = small_batch
    conf = _conf(2)
    want = _run_host(H, hb, conf, 2)
    L = _cabi.lib()
    pipe = L.raer_pipe_create(0)
    assert pipe, _cabi.last_error()
    try:
        got = []
        inflight = 0
        for _ in range(5):
            if inflight == 2:
                rows = _cabi.Rows()
                assert L.raer_pipe_collect(pipe, C.byref(rows)) == 0, _cabi.last_error()
                got.append(_rows_to_np(rows, 2))
                L.raer_rows_free(C.byref(rows))
                inflight -= 1
            rc = L.raer_pipe_submit(pipe, C.byref(conf), C.byref(hb), 0, H["ref"].data_ptr(), H["ref"].numel())
            assert rc == 0, _cabi.last_error()
            inflight += 1
        # a third batch in flight is refused
        assert L.raer_pipe_submit(pipe, C.byref(conf), C.byref(hb), 0, H["ref"].data_ptr(), H["ref"].numel()) != 0
        while inflight:
            rows = _cabi.Rows()
            assert L.raer_pipe_collect(pipe, C.byref(rows)) == 0, _cabi.last_error()
            got.append(_rows_to_np(rows, 2))
            L.raer_rows_free(C.byref(rows))
            inflight -= 1
        assert len(got) == 5
        for g in got:
            _same(g, want)
    finally:
        L.raer_pipe_destroy(pipe)


def test_tile_list_overflow_is_rerun(small_batch, monkeypatch):
    H, hb = small_batch
    conf = _conf(2)
    want = _run_host(H, hb, conf, 2)
    monkeypatch.setenv("RAER_TEST_LISTCAP", "1000")
    _same(_run_host(H, hb, conf, 2), want)


@pytest.mark.parametrize("kernel", ["lane", "fast", "generic"])
def test_row_buffers_smaller_than_the_rows_are_rerun(small_batch, monkeypatch, kernel):
    # the row buffers are sized from the reads, not from the genomic span; when the guess is too small the device
    # reports how many rows there are and the batch runs again (here the first guess is forced down to 100 rows)
    H, hb = small_batch
    conf = _conf(2, okv=0)
    if kernel != "lane":
        monkeypatch.setenv("RAER_KERNEL", kernel)
    want = _run_host(H, hb, conf, 2)
    assert want["pos"].size > 50_000
    monkeypatch.setenv("RAER_TEST_ROWCAP", "100")
    _same(_run_host(H, hb, conf, 2), want)
    # both guesses too small at once: tile lists and rows
    monkeypatch.setenv("RAER_TEST_LISTCAP", "1000")
    _same(_run_host(H, hb, conf, 2), want)


def test_coverage_beyond_16_bit_counters_falls_back(monkeypatch):
    # ~1.3e5 reads over 150 positions on one strand block: coverage above 65535 per (file, strand)
    H, hb = _host_batch(1, 250, 100_000, 5)
    conf = _conf(1, okv=0)
    fast = _run_host(H, hb, conf, 1)
    depth = fast["counts"][:, 0, 0] + fast["counts"][:, 0, 1]
    assert depth.max() > 65535, depth.max()
    monkeypatch.setenv("RAER_KERNEL", "generic")
    _same(fast, _run_host(H, hb, conf, 1))
