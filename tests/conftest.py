import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Everything compiled (CUDA engine, host layer, oracle).  Cheap when up to date."""
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "acc_genomics_b200", "libpairhmm_b200.so")) or \
       not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        g.build()
    return True


@pytest.fixture(scope="session")
def port(built):
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def reference(built):
    """The reference's own AVX code (oracle/_ref); None when it has not been built and cannot be (no /root/reference)."""
    import oracle
    return oracle.reference()


@pytest.fixture(scope="session")
def checker(port, reference):
    """The strongest checker available: the reference itself, else our restatement (bit-identical, see test_oracle)."""
    return reference or port


@pytest.fixture(scope="session")
def golden():
    from acc_genomics_b200.batch import Batch
    g = np.load(os.path.join(ROOT, "tests", "golden", "pairhmm_golden.npz"))
    names = sorted({k.split("/")[0] for k in g.files if not k.startswith("tables/")})
    cases = {}
    for n in names:
        b = Batch(*[g[f"{n}/{f}"] for f in ("read_off", "rs", "q", "i", "d", "c", "hap_off", "hap")])
        cases[n] = (b, g[f"{n}/raw_bits"], g[f"{n}/log10_bits"], g[f"{n}/mask"])
    tables = {k.split("/")[1]: g[k] for k in g.files if k.startswith("tables/")}
    return cases, tables


@pytest.fixture(scope="session")
def engine(built):
    from acc_genomics_b200.engine import PairHMMEngine
    e = PairHMMEngine(0)
    yield e
    e.close()


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    va = a.view(np.uint32 if a.dtype.itemsize == 4 else np.uint64)
    vb = b.view(np.uint32 if b.dtype.itemsize == 4 else np.uint64)
    bad = np.argwhere(va != vb)
    assert len(bad) == 0, f"{what}: {len(bad)}/{va.size} values differ, first at {bad[0]}: {a[tuple(bad[0])]!r} vs {b[tuple(bad[0])]!r}"
