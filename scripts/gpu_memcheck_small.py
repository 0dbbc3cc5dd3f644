"""Small inputs through the round-2 kernels, meant to run under `compute-sanitizer --tool memcheck` (development aid):
k-mer guided edit (grouped and one pair per warp, with fallback pairs and a starved pool), device-side shard packing, the wavefront kernel, the re-alignment kernel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers as ck  # noqa: E402
import torch  # noqa: E402
from bsalign_b200 import api, synth  # noqa: E402

ctx = api.Context(0)
rng = np.random.default_rng(3)


def related(n, qlen, p):
    q = rng.integers(0, 4, (n, qlen)).astype(np.uint8)
    t, tl = synth.mutate_batch(rng, q, p, p, p)
    out, o = [], 0
    for i in range(n):
        out.append((q[i].copy(), t[o + (i % 4) * 5:o + tl[i]].copy()))
        o += tl[i]
    return out


pairs = related(300, 300, .03) + related(40, 1000, .05) + related(60, 80, .1) + related(2, 6000, .03)
pairs += [(rng.integers(0, 4, 90).astype(np.uint8), rng.integers(0, 4, 70).astype(np.uint8)) for _ in range(10)]
core = rng.integers(0, 4, 1500).astype(np.uint8)
gap = np.concatenate([core[:200], rng.integers(0, 4, 900).astype(np.uint8), core[1200:]])
pairs += [(core, gap), (gap, core)] * 3
b = synth.PairBatch.from_lists(pairs)
exp, ecg, _ = ck.kmer_batch("oracle", b, 13, nthreads=8)
bad = 0
for env in ({}, {"BSB200_KMER_GROUP": "32"}, {"BSB200_KMER_GROUP": "1", "BSB200_KMER_POOL": "700000"}, {"BSB200_KMER_NOSMEM": "1", "BSB200_KMER_GROUP": "5"}):
    for k in ("BSB200_KMER_GROUP", "BSB200_KMER_POOL", "BSB200_KMER_NOSMEM"):
        os.environ.pop(k, None)
    os.environ.update(env)
    got = ctx.kmer_edit_batch(b, 13, dense=bool(env))
    ok = np.array_equal(got.results, exp) and all(np.array_equal(x, y) for x, y in zip(got.cigars(), ecg))
    print("kmer", env, "ok" if ok else "DIFFERS")
    bad += not ok
# device-side packing against the host packer
idx = rng.permutation(b.n)[:200].astype(np.uint64)
pb, nb = api.pack_pairs(b, idx)
src = torch.from_numpy(b.seqs).cuda()
dst = torch.zeros(max(nb, 1), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
api.pack_pairs_dev(ctx, src.data_ptr(), b, idx, dst.data_ptr())
ok = np.array_equal(dst.cpu().numpy()[:nb], pb.seqs[:nb])
print("pack_pairs_dev", "ok" if ok else "DIFFERS")
bad += not ok
# wavefront kernel, plain and split
mtx = synth.score_matrix(2, -6)
for qlen, n in ((400, 64), (3000, 6)):
    bb = synth.make_pairs(n, qlen, seed=44 + qlen)
    got = ctx.epi8_batch(bb, 0, 0, mtx, -3, -2, 0, 0)
    e2, c2, _ = ck.oracle_batch("epi8", bb, 0, 0, mtx, (-3, -2, 0, 0), nthreads=4)
    ok = np.array_equal(got.results, e2) and all(np.array_equal(x, y) for x, y in zip(got.cigars(), c2))
    print("wave", qlen, "ok" if ok else "DIFFERS")
    bad += not ok
# read re-alignment against the MSA profile: the reference's own records, with and without the byte matrices
import remsa_jobs as rj  # noqa: E402
jobs = rj.load_golden()[:12] + rj.load_golden()[-9:]
ms, out, mm = api.remsa_batch(ctx, jobs, want_matrices=True)
ms2, out2, _ = api.remsa_batch(ctx, jobs)
ok = all(rj.compare(j, mm[k][0], mm[k][1], ms[k]) is None and np.array_equal(ms[k], ms2[k]) for k, j in enumerate(jobs)) and np.array_equal(out, out2) and not out[:, 1].any()
print("remsa", "ok" if ok else "DIFFERS")
bad += not ok
ctx.close()
sys.exit(1 if bad else 0)
