import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

EXT = os.path.join(ROOT, "tests", "golden", "extdata")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def ext():
    class P:
        bam1 = os.path.join(EXT, "SRR5564269_Aligned.sortedByCoord.out.md.bam")
        bam2 = os.path.join(EXT, "SRR5564277_Aligned.sortedByCoord.out.md.bam")
        fasta = os.path.join(EXT, "human.fasta")
        bed = os.path.join(EXT, "regions.bed")
        scbam = os.path.join(EXT, "5k_neuron_mouse_possort.bam")
        mouse_fa = os.path.join(EXT, "mouse_tiny.fasta")
        sc_sites_bed = os.path.join(EXT, "5k_neuron_sites.bed.gz")

    return P
