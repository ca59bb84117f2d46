"""Host logic: wire format and the synthetic generator."""
import hashlib

import numpy as np

from acc_genomics_b200 import synth
from acc_genomics_b200.batch import Batch, deserialize, serialize_haps, serialize_reads


def test_wire_format_sizes_match_the_reference():
    """config 1 serialises to 65 156 B of reads and 12 932 B of haplotypes
    (4 + n*(4 + 5*len) and 4 + n*(4 + len), PairHMMHostInterface.cpp:175-207)."""
    b = synth.config(1)[0]
    assert len(serialize_reads(b)) == 4 + 128 * (4 + 5 * 101) == 65156
    assert len(serialize_haps(b)) == 4 + 32 * (4 + 400) == 12932


def test_wire_format_round_trip():
    b = synth.config(5, scale=0.0004)[0]
    r = deserialize(serialize_reads(b), serialize_haps(b))
    for f in ("read_off", "rs", "q", "i", "d", "c", "hap_off", "hap"):
        assert np.array_equal(getattr(b, f), getattr(r, f)), f
    blob = serialize_reads(b)
    n0 = int(b.read_lens[0])
    assert np.frombuffer(blob, dtype=np.int32, count=1)[0] == b.num_read
    assert np.frombuffer(blob, dtype=np.int32, count=1, offset=4)[0] == n0
    assert blob[8:8 + n0] == b.rs[:n0].tobytes() and blob[8 + n0:8 + 2 * n0] == b.q[:n0].tobytes()


def test_generator_is_deterministic_and_shaped():
    a, b = synth.config(2, scale=0.05)[0], synth.config(2, scale=0.05)[0]
    h = lambda x: hashlib.sha256(b"".join(getattr(x, f).tobytes() for f in ("read_off", "rs", "q", "i", "d", "c", "hap_off", "hap"))).hexdigest()
    assert h(a) == h(b)
    assert h(a) != h(synth.config(2, seed=99, scale=0.05)[0])
    full = synth.config(1)[0]
    assert full.num_read == 128 and full.num_hap == 32 and full.num_cells == 165478400
    c4 = synth.config(4, scale=0.1)[0]
    assert (c4.read_lens == 250).all() and c4.hap_lens.min() >= 1000 and c4.hap_lens.max() <= 2048
    c5 = synth.config(5, scale=0.002)
    assert len(c5) == 5 and all(r.num_read == 100 and r.num_hap == 40 for r in c5)
    assert min(r.read_lens.min() for r in c5) >= 60 and max(r.read_lens.max() for r in c5) == 151
    assert set(np.unique(full.rs)) <= set(b"ACGTN")
