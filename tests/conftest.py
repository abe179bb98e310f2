import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def meshes():
    """The reference's assets as its own importer parses them (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(HERE, "golden", "meshes.npz"))
    names = sorted({k[:-2] for k in z.files})
    return {m: (z[m + "_v"], z[m + "_t"]) for m in names}


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(HERE, "golden", "ref_digests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_large():
    """Digests of the reference's OpenMP JFA at 512^3 / 1024^3 (tests/golden/make_golden_large.py)."""
    with open(os.path.join(HERE, "golden", "ref_digests_large.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from checkers import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from checkers import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libvpref.so not built (needs /root/reference)")
    return Reference()
