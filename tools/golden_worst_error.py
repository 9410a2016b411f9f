import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import golden_util
from difffr_b200.cabi import Context
f = lambda **k: Context(device=0, **k)
for name in golden_util.PAPER_CASES:
    print(name, "worst rel err (CUDA vs reference golden)", golden_util.replay_paper_and_compare(f, name, 1e-6, 1e-4))
for name in golden_util.CASES:
    print(name, golden_util.replay_and_compare(f, name, 1e-6, 1e-4))
