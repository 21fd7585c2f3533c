"""The R-facing boundary (raer_b200/csrc/r_shim.c) compiled and driven without R: tests/r_stub/ implements the part of
R's C API the shim uses, plus a harness that .Calls the registered routines. Checked here: the registration table the
reference's NAMESPACE relies on (src/init.c:12-24), the argument checks with the reference's messages
(src/plp.c:998-1055), the result layout of .pileup (src/plp_data.c:285-361), the interrupt hook
(src/plp.c:925-931, src/plp_utils.c:7-12), ownership of the native result when R fails to allocate, and the two
additional routines .pileup_matrix / .pileup_aei."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from raer_b200 import FilterParam, _cabi
from raer_b200.glue import Sites, prepare_pileup_call

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NILSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP = 0, 10, 13, 14, 16, 19


@pytest.fixture(scope="module")
def R(tmp_path_factory):
    """the shim + the stub runtime as one shared object, linked against libraer_b200.so"""
    d = tmp_path_factory.mktemp("rshim")
    so = str(d / "raer_shim_test.so")
    stub = os.path.join(ROOT, "tests", "r_stub")
    lib = os.path.join(ROOT, "raer_b200")
    _cabi.lib()  # builds nothing, but fails loudly if the product library is missing
    subprocess.run(["gcc", "-O1", "-g", "-fPIC", "-shared", "-std=gnu11", "-Wall", "-Werror", "-Wno-unused-function",
                    "-I", stub, "-I", os.path.join(ROOT, "include"), "-o", so,
                    os.path.join(lib, "csrc", "r_shim.c"), os.path.join(stub, "r_stub.c"),
                    "-L", lib, "-lraer_b200", f"-Wl,-rpath,{lib}"], check=True)
    L = C.CDLL(so)
    P = C.c_void_p
    for nm, res, args in [
        ("rstub_init", C.c_int, []), ("rstub_nil", P, []), ("rstub_int", P, [C.c_int, C.POINTER(C.c_int)]),
        ("rstub_lgl", P, [C.c_int, C.POINTER(C.c_int)]), ("rstub_real", P, [C.c_int, C.POINTER(C.c_double)]),
        ("rstub_str", P, [C.c_int, C.POINTER(C.c_char_p)]), ("rstub_list", P, [C.c_int]),
        ("rstub_list_set", None, [P, C.c_int, P]), ("rstub_set_dim", None, [P, C.c_int, C.c_int]),
        ("rstub_type", C.c_int, [P]), ("rstub_len", C.c_long, [P]), ("rstub_data", P, [P]),
        ("rstub_chr", C.c_char_p, [P, C.c_long]), ("rstub_elt", P, [P, C.c_long]), ("rstub_names", P, [P]),
        ("rstub_dim", P, [P]), ("rstub_message", C.c_char_p, []), ("rstub_interrupt_at_check", None, [C.c_long]),
        ("rstub_checks", C.c_long, []), ("rstub_fail_alloc_at", None, [C.c_long]), ("rstub_run_finalizers", C.c_long, []),
        ("rstub_call", C.c_int, [C.c_char_p, C.c_int, C.POINTER(P), C.POINTER(P)]),
        ("rstub_registered", C.c_int, [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int)]),
    ]:
        f = getattr(L, nm)
        f.restype, f.argtypes = res, args
    assert L.rstub_init() == 1
    return L


class Conv:
    """Python values <-> stub SEXPs"""

    def __init__(self, L):
        self.L = L

    def ints(self, xs):
        xs = [int(x) for x in xs]
        return self.L.rstub_int(len(xs), (C.c_int * max(len(xs), 1))(*xs))

    def lgls(self, xs):
        xs = [int(bool(x)) for x in xs]
        return self.L.rstub_lgl(len(xs), (C.c_int * max(len(xs), 1))(*xs))

    def reals(self, xs):
        xs = [float(x) for x in xs]
        return self.L.rstub_real(len(xs), (C.c_double * max(len(xs), 1))(*xs))

    def strs(self, xs):
        xs = [x.encode() for x in xs]
        return self.L.rstub_str(len(xs), (C.c_char_p * max(len(xs), 1))(*xs))

    def lst(self, items):
        l = self.L.rstub_list(len(items))
        for i, it in enumerate(items):
            self.L.rstub_list_set(l, i, it)
        return l

    def pileup_args(self, pc):
        """the 13 arguments of .Call(".pileup", ...) for a prepared call (R/pileup.R:183-198)"""
        lst = self.L.rstub_nil()
        if pc.lst is not None and len(pc.lst[0]) > 0:
            lst = self.lst([self.strs(pc.lst[0]), self.ints(pc.lst[1]), self.ints(pc.lst[2])])
        return [self.strs(pc.bampaths), self.strs(pc.indexes), self.ints([pc.n]), self.strs([pc.fapath]),
                self.strs([pc.region] if pc.region else []), lst, self.ints(pc.int_args), self.reals(pc.dbl_args),
                self.lgls(pc.lgl_args), self.ints(pc.libtype), self.lgls(pc.only_keep_variants), self.ints(pc.min_mapq),
                self.strs([pc.umi] if pc.umi else [])]

    def call(self, name, args):
        arr = (C.c_void_p * len(args))(*args)
        out = C.c_void_p()
        rc = self.L.rstub_call(name.encode(), len(args), arr, C.byref(out))
        return rc, out.value, self.L.rstub_message().decode()

    def value(self, x):
        """SEXP -> numpy / list (names kept as dict keys when present)"""
        L = self.L
        t, n = L.rstub_type(x), L.rstub_len(x)
        if t == NILSXP:
            return None
        if t in (INTSXP, LGLSXP):
            v = np.ctypeslib.as_array(C.cast(L.rstub_data(x), C.POINTER(C.c_int)), (max(n, 1),))[:n].copy()
        elif t == REALSXP:
            v = np.ctypeslib.as_array(C.cast(L.rstub_data(x), C.POINTER(C.c_double)), (max(n, 1),))[:n].copy()
        elif t == STRSXP:
            v = np.array([L.rstub_chr(x, i).decode() for i in range(n)], dtype=object)
        elif t == VECSXP:
            items = [self.value(L.rstub_elt(x, i)) for i in range(n)]
            nm = L.rstub_names(x)
            if L.rstub_type(nm) == STRSXP:
                return {L.rstub_chr(nm, i).decode(): items[i] for i in range(n)}
            return items
        else:
            raise AssertionError(f"unexpected SEXP type {t}")
        dim = L.rstub_dim(x)
        if L.rstub_type(dim) == INTSXP:
            d = self.value(dim)
            v = v.reshape(tuple(int(k) for k in d), order="F")
        return v


@pytest.fixture(scope="module")
def cv(R):
    return Conv(R)


def test_registration_table_matches_the_reference(R):
    got = {}
    name, arity = C.c_char_p(), C.c_int()
    i = 0
    while R.rstub_registered(i, C.byref(name), C.byref(arity)):
        got[name.value.decode()] = arity.value
        i += 1
    # src/init.c:12-24: the four routines NAMESPACE's useDynLib(raer, .registration = TRUE) resolves
    for nm, ar in {".get_region": 1, ".fisher_exact": 1, ".pileup": 13, ".scpileup": 14}.items():
        assert got.get(nm) == ar, (nm, got)
    assert got.get(".pileup_matrix") == 13 and got.get(".pileup_aei") == 13


def test_get_region_and_fisher_exact_through_the_shim(R, cv):
    rc, out, msg = cv.call(".get_region", [cv.strs(["chr1:1,000-2,000"])])
    assert rc == 0, msg
    v = cv.value(out)
    assert list(v["chrom"]) == ["chr1"] and int(v["start"][0]) == 999 and int(v["end"][0]) == 2000
    rc, out, msg = cv.call(".get_region", [cv.ints([1])])
    assert rc == 1 and msg == "'region' must be character"
    m = cv.ints([3, 1, 1, 3, 10, 0, 0, 10])
    R.rstub_set_dim(m, 4, 2)
    rc, out, msg = cv.call(".fisher_exact", [m])
    assert rc == 0, msg
    want = _cabi.fisher_exact(np.array([[3, 1, 1, 3], [10, 0, 0, 10]], dtype=np.int32).T)
    assert np.allclose(cv.value(out), want, rtol=1e-12)
    rc, out, msg = cv.call(".fisher_exact", [cv.ints([1, 2, 3])])
    assert rc == 1 and msg == "'mat' must be matrix with 4 rows"


def test_argument_checks_raise_the_reference_messages(ext, cv):
    pc = prepare_pileup_call([ext.bam1, ext.bam2], ext.fasta)
    good = cv.pileup_args(pc)
    cases = [
        (2, cv.reals([2.0]), "'n' must be integer(1)"),
        (0, cv.strs([ext.bam1]), "'bampaths' must be character vector equal in length to number of bam files"),
        (3, cv.strs([]), "'fapath' must be character(1)"),
        (6, cv.ints([1, 2, 3]), "'int_args' must be integer of length 14"),
        (7, cv.reals([0.0]), "'dbl_args' must be numeric of length 5"),
        (8, cv.lgls([True]), "'lgl_args' must be logical of length 2"),
        (9, cv.ints([1]), "'lib_type' must be integer of same length as bamfiles"),
        (10, cv.lgls([False]), "'only_keep_variants' must be logical of same length as bamfiles"),
        (11, cv.ints([0]), "'min_mapQ' must be integer of same length as bamfiles"),
        (12, cv.strs(["UB", "CB"]), "'umi' must be character of length 0 or 1"),
    ]
    for idx, bad, want in cases:
        for routine in (".pileup", ".pileup_matrix", ".pileup_aei"):
            a = list(good)
            a[idx] = bad
            rc, _, msg = cv.call(routine, a)
            assert rc == 1 and msg == want, (routine, idx, msg)
    # arity is part of the registration: a 12-argument call does not resolve
    assert cv.call(".pileup", good[:12])[0] == -1


def test_no_device_is_an_r_error_not_a_crash(ext, cv):
    import torch

    if torch.cuda.is_available():
        pytest.skip("device present")
    pc = prepare_pileup_call(ext.bam1, ext.fasta)
    rc, _, msg = cv.call(".pileup", cv.pileup_args(pc))
    assert rc == 1 and msg.startswith("[raer internal] error detected during pileup") and "no usable CUDA device" in msg


def test_poll_callback_stops_the_ingest(ext):
    """the library side of the interrupt hook, host only: raer_ingest_summary streams like raer_pileup()"""
    L = _cabi.lib()
    L.raer_ingest_summary.argtypes = [C.POINTER(_cabi.PileupArgs), C.c_int64, C.POINTER(C.c_int64)]
    pc = prepare_pileup_call(ext.bam1, ext.fasta)
    a = _cabi.PileupArgs()
    keep = [_cabi._strs(pc.bampaths), _cabi._strs(pc.indexes)]
    a.nbam, a.bampaths, a.indexes, a.fapath = pc.n, keep[0], keep[1], pc.fapath.encode()
    a.int_args = (C.c_int32 * 14)(*pc.int_args)
    a.dbl_args = (C.c_double * 5)(*pc.dbl_args)
    a.lgl_args = (C.c_int32 * 2)(*[int(x) for x in pc.lgl_args])
    keep += [_cabi._ints(pc.libtype), _cabi._ints([int(x) for x in pc.only_keep_variants]), _cabi._ints(pc.min_mapq)]
    a.libtype, a.only_keep_variants, a.min_mapq = keep[-3:]
    calls = []

    @_cabi.POLL_FN
    def poll(_ctx):
        calls.append(1)
        return 1 if len(calls) >= 3 else 0

    a.poll = C.cast(poll, C.c_void_p)
    out = (C.c_int64 * 16)()
    rc = L.raer_ingest_summary(C.byref(a), 40, out)  # 40 reads per batch: many batches
    assert rc == -8 and len(calls) == 3 and "interrupted" in _cabi.last_error()  # RAER_ERR_INTERRUPT
    a.poll = None
    assert L.raer_ingest_summary(C.byref(a), 40, out) == 0


# ------------------------------------------------------------------------------------------ with a device
def _columns_equal(lst, cols, f):
    d = cols["files"][f]
    assert list(lst["seqname"]) == list(cols["seqname"])
    assert np.array_equal(lst["pos"], cols["pos"]) and list(lst["strand"]) == list(cols["strand"])
    assert list(lst["REF"]) == list(cols["REF"]) and list(lst["ALT"]) == list(d["ALT"])
    for nm in ("nRef", "nAlt", "nA", "nT", "nC", "nG", "nN", "nX"):
        assert np.array_equal(lst[nm], d[nm]), nm


@pytest.mark.gpu
@pytest.mark.parametrize("kw,sites", [(dict(), False), (dict(library_type="fr-first-strand", only_keep_variants=True), False),
                                      (dict(min_depth=2), True)])
def test_pileup_through_the_shim_equals_the_c_abi(ext, cv, kw, sites):
    st = Sites.from_bed(ext.bed) if sites else None
    pc = prepare_pileup_call([ext.bam1, ext.bam2], ext.fasta, st, None, None, FilterParam(**kw))
    want = _cabi.pileup(pc)
    rc, out, msg = cv.call(".pileup", cv.pileup_args(pc))
    assert rc == 0, msg
    v = cv.value(out)
    assert len(v) == 3 and list(v[0]) == ["rpbz", "vdb", "sor"]  # src/plp_data.c:347-361
    assert list(v[1]) == ["seqname", "pos", "strand", "REF", "ALT", "nRef", "nAlt", "nA", "nT", "nC", "nG", "nN", "nX"]
    for k in ("rpbz", "vdb", "sor"):
        assert np.array_equal(v[0][k], want[k], equal_nan=True)
    assert want["nrows"] > 50
    for f in range(2):
        _columns_equal(v[1 + f], want, f)
    # the matrix routine (8f-1): rowRanges columns once, one rows x files matrix per assay
    rc, out, msg = cv.call(".pileup_matrix", cv.pileup_args(pc))
    assert rc == 0, msg
    m = cv.value(out)
    assert list(m["seqname"]) == list(want["seqname"]) and np.array_equal(m["pos"], want["pos"])
    assert m["nRef"].shape == (want["nrows"], 2) and m["ALT"].shape == (want["nrows"], 2)
    for f in range(2):
        for nm in ("nRef", "nAlt", "nA", "nT", "nC", "nG", "nN", "nX"):
            assert np.array_equal(m[nm][:, f], want["files"][f][nm]), nm
        assert list(m["ALT"][:, f]) == list(want["files"][f]["ALT"])


@pytest.mark.gpu
def test_pileup_aei_through_the_shim(ext, cv):
    pc = prepare_pileup_call([ext.bam1, ext.bam2], ext.fasta, None, None, None,
                             FilterParam(library_type="fr-first-strand", min_depth=1))
    want = _cabi.pileup(pc, aei=True)["aei"]
    rc, out, msg = cv.call(".pileup_aei", cv.pileup_args(pc))
    assert rc == 0, msg
    got = cv.value(out)  # [REF, base, file] column-major
    assert got.shape == (4, 4, 2) and want.sum() > 1000
    for f in range(2):
        assert np.array_equal(got[:, :, f].astype(np.int64), want[f])


@pytest.mark.gpu
def test_interrupt_unwinds_cleanly(ext, cv, R):
    pc = prepare_pileup_call([ext.bam1, ext.bam2], ext.fasta)
    R.rstub_interrupt_at_check(1)  # the first R_CheckUserInterrupt() of the call finds a pending interrupt
    rc, _, msg = cv.call(".pileup", cv.pileup_args(pc))
    assert rc == 2 and msg == "interrupt" and R.rstub_checks() >= 1
    assert R.rstub_run_finalizers() == 0  # nothing native was left behind
    rc, out, msg = cv.call(".pileup", cv.pileup_args(pc))  # and the next call works
    assert rc == 0, msg


@pytest.mark.gpu
def test_native_result_is_released_when_r_cannot_allocate(ext, cv, R):
    pc = prepare_pileup_call(ext.bam1, ext.fasta)
    args = cv.pileup_args(pc)
    R.rstub_run_finalizers()
    R.rstub_fail_alloc_at(1)  # the first allocVector() after raer_pileup() returned
    rc, _, msg = cv.call(".pileup", args)
    assert rc == 1 and msg.startswith("cannot allocate")
    assert R.rstub_run_finalizers() == 1  # the external pointer's finalizer released the native result
