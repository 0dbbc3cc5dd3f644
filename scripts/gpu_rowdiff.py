"""Development aid: compare the GPU forward trace of one pair row by row with the oracle's dump."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import checkers as ck
from bsalign_b200 import api, synth

def rowdiff(ctx, q, t, mode, bw_req, mtx, gaps, maxshow=3):
    L = api.lib()
    L.bsb200_debug_trace.restype = ctypes.c_int64
    L.bsb200_debug_trace.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    b = synth.PairBatch.from_lists([(q, t)])
    rb = ctx.upload("epi8", b, mode, bw_req, mtx, gaps)
    rb.run()
    buf = np.zeros(64 << 20, dtype=np.uint8)
    bw = ctypes.c_uint32(0); pw = ctypes.c_int(0); ubias = ctypes.c_int(0)
    nbytes = L.bsb200_debug_trace(ctx._h, rb._h, 0, buf.ctypes.data, buf.nbytes, ctypes.byref(bw), ctypes.byref(pw), ctypes.byref(ubias))
    assert nbytes > 0, nbytes
    bw, pw = bw.value, pw.value
    tlen = len(t)
    W = bw // 16
    IB = (W + 7) // 8 * 128
    AB = ((W + 31) // 32 - 1) * 64 if W > 64 else 0
    RS = IB * (pw + 1) + AB
    rows = buf[:RS * (tlen + 1)].reshape(tlen + 1, RS)[:, :IB * (pw + 1)].reshape(tlen + 1, pw + 1, IB).view(np.int8)
    pp = np.arange(bw); jj = pp // W; ii = pp % W
    idx = (ii >> 3) * 128 + (jj >> 1) * 16 + (ii & 7) * 2 + (jj & 1)
    meta = buf[RS * (tlen + 1):RS * (tlen + 1) + 80 * (tlen + 1)].view(np.int32).reshape(tlen + 1, 20)
    res, begs, ub, u, e, qq = ck.rows_dump(ck.oracle(), "bso_epi8_pairwise_ex", q, t, mode, bw_req, mtx, gaps,
                                           extra_args=(None, ctypes.c_uint32(0), None))
    got = rb.fetch()
    print("gpu result", got.results[0], "status", got.status[0]); print("exp result", res)
    shown = 0
    for y in range(tlen):
        gb = int(meta[y + 1, 17]); gub = meta[y + 1, :17]
        gu = rows[y + 1, 0][idx]
        if ubias.value:
            gu = (gu.view(np.uint8).astype(np.int16) - ubias.value).astype(np.int8)
        ge = rows[y + 1, 1][idx] if pw >= 1 else None
        gq = rows[y + 1, 2][idx] if pw == 2 else None
        ok = gb == begs[y] and np.array_equal(gub, ub[y]) and np.array_equal(gu, u[y]) and (pw < 1 or np.array_equal(ge, e[y])) and (pw < 2 or np.array_equal(gq, qq[y]))
        if not ok:
            print("row", y, "beg gpu/exp", gb, begs[y])
            print(" ub gpu", gub); print(" ub exp", ub[y])
            d = np.nonzero(gu != u[y])[0]; print(" u diff at", d[:20], "gpu", gu[d[:10]], "exp", u[y][d[:10]])
            if pw >= 1:
                d = np.nonzero(ge != e[y])[0]; print(" e diff at", d[:20], "gpu", ge[d[:10]], "exp", e[y][d[:10]])
            shown += 1
            if shown >= maxshow: break
    if shown == 0: print("all", tlen, "rows identical")
    rb.free()

if __name__ == "__main__":
    ctx = api.Context(0)
    m = synth.score_matrix(2, -6)
    b = synth.make_pairs(1, 300, seed=0)
    rowdiff(ctx, b.query(0), b.target(0), 0, 128, m, (0, -2, 0, 0))
    b = synth.make_pairs(1, 100, seed=5)
    rowdiff(ctx, b.query(0), b.target(0), 0, 32, m, (-3, -2, 0, 0))
    rowdiff(ctx, b.query(0), b.target(0), 1, 0, m, (-3, -2, 0, 0))
    b = synth.make_pairs(1, 400, seed=6)
    rowdiff(ctx, b.query(0), b.target(0), 1, 64, m, (-3, -2, -8, -1))
    rowdiff(ctx, b.query(0), b.target(0)[:150], 0, 64, m, (-3, -2, 0, 0))
    rowdiff(ctx, b.query(0), b.target(0), 0, 48, synth.score_matrix(30, -40), (-40, -20, 0, 0))
