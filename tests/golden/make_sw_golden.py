"""Mint tests/golden/sw_golden.json from the REFERENCE's own Smith-Waterman (oracle/_ref/libsw_ref.so, compiled from
/root/reference/htc-sw by oracle/Makefile).  Run in the build container, where /root/reference exists:
    python tests/golden/make_sw_golden.py
Every case records the inputs, the overhang strategy and what runSWOnePairBT_fp_avx2 returned (alignment offset,
CIGAR); Falcon's SWPairwiseAlignmentOneBatch is asserted to agree while minting."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
from acc_genomics_b200 import sw

ref = oracle.sw_reference()
assert ref is not None, "needs /root/reference and an AVX2 host"
cases = []
pairs = sw.haplotype_pairs(101, 24, ref_len=(40, 160), per_ref=4) + sw.haplotype_pairs(102, 8, ref_len=(300, 420), per_ref=2) + \
        [(b"A", b"A"), (b"A", b"C"), (b"ACGT", b"T"), (b"G", b"ACGTACGT"), (b"ACGTACGTAC", b"ACGTTTACGTAC"), (b"ACGTACGTACGTACGT", b"ACGTACGTACGT"),
         (b"TTTTTTTTTT", b"TTTTT"), (b"ACACACACACAC", b"ACACACAC"), (b"GATTACAGATTACA", b"CATTAGAGATTTACA")]
for r, a in pairs:
    for st in range(4):
        off, cig = ref.align(r, a, st)
        rc, off2, cig2 = ref.align_falcon(r, a, st, 1)
        assert rc != 0 or (off2, cig2) == (off, cig), (r, a, st)
        cases.append({"ref": r.decode(), "alt": a.decode(), "strategy": st, "offset": off, "cigar": cig})
# non-default weights exercise the parameters of runSWOnePairBT (Falcon's code has them compiled in)
for r, a in pairs[:6]:
    for w in ((100, -50, -100, -20), (25, -50, -110, -6)):
        off, cig = ref.align(r, a, 0, w)
        cases.append({"ref": r.decode(), "alt": a.decode(), "strategy": 0, "weights": list(w), "offset": off, "cigar": cig})
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sw_golden.json")
with open(path, "w") as f:
    json.dump(cases, f, separators=(",", ":"))
print(len(cases), "cases ->", path, os.path.getsize(path), "bytes")
