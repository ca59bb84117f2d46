import sys, os
sys.path.insert(0, os.getcwd())
from acc_genomics_b200 import synth
from acc_genomics_b200.engine import PairHMMEngine
eng = PairHMMEngine(0)
for kv in os.environ.get("PMM_OPTS", "").split():
    eng.set_option(*kv.split("="))
b = synth.config(int(os.environ.get("CFG", "2")))
eng.stage(b)
for _ in range(3): eng.launch()
eng.sync(); print(eng.stats())
