"""The ingest's own raw-DEFLATE decoder (raer_b200/csrc/raer_inflate.cpp) against zlib: random payloads of six kinds,
compression levels 0-9 and every zlib strategy (stored, fixed and dynamic blocks), exact output, no write outside the
output window, no crash on corrupted input. Also: a BAM read through the decoder equals the same BAM read through
zlib (RAER_ZLIB_INFLATE)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_inflate_matches_zlib(tmp_path):
    exe = str(tmp_path / "inflate_check")
    src = os.path.join(ROOT, "raer_b200", "csrc")
    subprocess.run(["g++", "-O2", "-I", src, os.path.join(ROOT, "tests", "inflate_check.cpp"),
                    os.path.join(src, "raer_inflate.cpp"), "-lz", "-o", exe], check=True)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "1500 streams, 0 mismatches, 0 refused" in p.stdout, p.stdout + p.stderr


def test_bam_through_own_inflate_equals_zlib(ext):
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from test_cabi_host import _ingest\n"
        "from raer_b200.glue import prepare_pileup_call\n"
        "print(sorted(_ingest(prepare_pileup_call([%r, %r], %r)).items()))\n"
    ) % (ROOT, os.path.join(ROOT, "tests"), ext.bam1, ext.bam2, ext.fasta)
    outs = []
    for zl in (False, True):
        env = dict(os.environ)
        env.pop("RAER_ZLIB_INFLATE", None)
        if zl:
            env["RAER_ZLIB_INFLATE"] = "1"
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
        assert p.returncode == 0, p.stderr
        outs.append(p.stdout)
    assert outs[0] == outs[1] and "records" in outs[0]


def test_qname_table_matches_unordered_map(tmp_path):
    """The parser's allocation-free mate table (QnameTable, raer_host.h) against std::unordered_map."""
    exe = str(tmp_path / "qname_table_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "raer_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "qname_table_check.cpp"), "-o", exe], check=True)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0 and "0 differences" in p.stdout, p.stdout + p.stderr
