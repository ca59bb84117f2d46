import sys, os
sys.path.insert(0, os.getcwd())
from acc_genomics_b200 import sw
al = sw.SmithWaterman(0)
pairs = sw.haplotype_pairs(1, 260, ref_len=(380, 420), per_ref=260)
for _ in range(3): al.align(pairs, 0)
print(al.stats())
