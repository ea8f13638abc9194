import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# the GPU tests work on small batches: let the banded fill (k_band_*) run at any size so that they exercise it
# (the library's default only uses it for launches with 16,384 or more reads to align: treat_b200_set_band_min_reads)
os.environ.setdefault("TREAT_B200_BAND_MIN", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fx():
    """Frozen reference inputs/expectations (tests/golden/make_golden.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "fixtures.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def examples(fx, tmp_path_factory):
    """Materialise the reference's examples/*.fa from the fixture file; returns {name: path}."""
    d = tmp_path_factory.mktemp("examples")
    out = {}
    for name, text in fx["fasta"].items():
        p = d / name
        p.write_bytes(text.encode("latin-1"))
        out[name] = str(p)
    return out
