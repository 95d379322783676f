import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

SMALL_GENOME = {"1": 6_010_000, "2": 4_800_000, "10": 3_333_333, "X": 5_000_001, "Y": 2_000_000, "M": 16571}
CHROMS = ["#", "X"]
SORTED_SMALL = ["1", "2", "10", "X"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_shim
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this machine")
    for item in items:
        if "reference" in item.keywords and not ref_shim.available():
            item.add_marker(skip_ref)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=True)


def unflatten(npz, prefix):
    """Inverse of oracle.make_golden.flatten for one top-level prefix."""
    out = {}
    for k in npz.files:
        parts = k.split("|")
        if parts[0] != prefix:
            continue
        d = out
        for p in parts[1:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = npz[k]
    return out


@pytest.fixture
def small_genome_file(tmp_path):
    from hichap_master_b200 import synth
    return synth.write_genome_size(str(tmp_path / "genomeSize"), SMALL_GENOME)


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
