"""One k-mer guided edit call on a synthetic configs[3]-shaped batch: the process `ncu -k regex:kmer_edit_kernel` is pointed at
(profiles/ncu_full_kmer_edit_*).  Usage: python scripts/kmer_profile_run.py [pairs] [qlen] [k]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsalign_b200 import api, synth  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
qlen = int(sys.argv[2]) if len(sys.argv) > 2 else 300
k = int(sys.argv[3]) if len(sys.argv) > 3 else 13
os.environ["BSB200_NOPIPE"] = "1"   # one launch over the whole batch
ctx = api.Context(0)
batch = synth.make_pairs(pairs, qlen, seed=5, p_sub=0.02, p_ins=0.02, p_del=0.02)
for _ in range(2):
    ctx.kmer_edit_batch(batch, k, dense=True)
print(ctx.timing())
