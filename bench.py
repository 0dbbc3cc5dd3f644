#!/usr/bin/env python
"""bench.py -- GCUPS of the banded striped DP hot path on N B200s (contract: see the task's section 4).

One "step" = one pass of the hot path over one batch of synthetic pairs.  Default workload is
BASELINE.json configs[1] ("c2": 100k pairs, 1 kb x 1 kb, global, 8-bit affine, full band) per GPU;
with N > 1 the ONE batch lives on rank 0 and is sharded: lengths broadcast, compact shard arenas scattered over
NVLink (NCCL), every rank aligns its shard, records + dense cigars gathered to rank 0 ("scaling": "strong";
bsalign_b200/shard.py).

  value : whole-job GCUPS with inputs resident in HBM (kernels only, CUDA events on the library stream)
  e2e   : same metric through the host-buffer C-ABI call (pinned host inputs -> H2D -> kernels -> D2H)
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md section 6

--impl reference times the reference's own CPU implementation (oracle/_ref/libbsref.so, the unmodified
headers; falls back to the oracle port when that was not built) on all host cores over a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

WORKLOADS = {
    # name: (kind, pairs per GPU, qlen, error model, mode, bandwidth, description)
    "c2": dict(kind="epi8", pairs=100000, qlen=1000, err=(0.03, 0.03, 0.04), mode=0, bandwidth=0,
               desc="BASELINE configs[1]: 100k pairs 1kb x 1kb, global 8-bit affine, bandwidth=0 (full)"),
    "c3": dict(kind="epi8", pairs=10000, qlen=10000, err="ont12", mode=1, bandwidth=512,
               desc="BASELINE configs[2]: 10k pairs 10kb x 10kb ONT-like, overlap, band 512"),
    "c4": dict(kind="edit", pairs=1000000, qlen=300, err=(0.02, 0.02, 0.02), mode=0, bandwidth=64,
               desc="BASELINE configs[3]: 1M pairs 300bp x 300bp, 2-bit edit, band 64"),
    # g10k: `pairs` is the fallback; on a GPU main() takes as many pairs as one wave of the traceback store seats
    "g10k": dict(kind="epi8", pairs=592, qlen=10000, err=(0.03, 0.03, 0.04), mode=0, bandwidth=0,
                 desc="north_star target: 10kb x 10kb global, full band"),
    # POA: `pairs` = MSA jobs per GPU stepped in lock-step; one step = one sweep round (every job aligns its next read
    # against the graph of its previous nrec+1 = 21 reads, bspoa.h:2636-2642); seqcore = 40 makes 39 such rounds per 64-read job
    "c5": dict(kind="poa", pairs=1000, qlen=15000, err=(0.03, 0.03, 0.04), mode=1, bandwidth=128, distinct=16,
               desc="BASELINE configs[4]: BSPOA 64 reads x 15 kb (DEFAULT_BSPOA_PAR), 1k MSA jobs in lock-step; step = one "
                    "read-vs-graph sweep round (graph of nrec+1 = 21 reads) over all jobs; DEFAULT_BSPOA_PAR (seqcore 40) runs 39 such rounds per 64-read job"),
}
MATRIX = (2, -6)
GAPS = (-3, -2, 0, 0)


def make_batch(w, seed, pairs):
    from bsalign_b200 import synth
    err = synth.ont_like(0.12) if w["err"] == "ont12" else w["err"]
    return synth.make_pairs(pairs, w["qlen"], seed, *err)


def nominal_cells(w, batch):
    """bw_eff * tlen per pair."""
    q = batch.qlen.astype(np.int64)
    t = batch.tlen.astype(np.int64)
    if w["kind"] == "epi8":
        bw = np.where(w["bandwidth"] == 0, q, w["bandwidth"])
        bw = (bw + 15) // 16 * 16
        bw = np.minimum(bw, (q + 15) // 16 * 16)
    else:  # bsalign.h:1055-1067
        q64 = (q + 63) // 64 * 64
        if w["mode"] in (1, 2):
            bw = q64
        else:
            bw = np.full_like(q, (w["bandwidth"] + 63) // 64 * 64)
            bw = np.where((bw == 0) | (bw > q), q64, bw)
            need = (q + t - 1) // np.maximum(t, 1) + 1
            bw = np.where((bw < q) & (bw < need), (need + 63) // 64 * 64, bw)
    return int((bw * t).sum())


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Start of the timed region: nvidia-smi is started earlier (before the warm-up steps) because it needs a few hundred
        ms to come up, longer with eight ranks starting one each; only samples from here on are used."""
        self.t_begin = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        if not any(t >= getattr(self, "t_begin", 0.0) for t, _ in self.lines):
            time.sleep(0.45)   # region shorter than the sampling period: take the samples right behind it (clocks ramp down after seconds)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        t0 = getattr(self, "t_begin", 0.0)
        inside = [ln for t, ln in self.lines if t0 <= t <= t_end + 0.25]
        if not inside:
            inside = [ln for t, ln in self.lines if t >= t0][:2] or [ln for _, ln in self.lines[-1:]]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(w, batch, nthreads, repeat=1, want_cigars=False):
    """Reference CPU implementation (or the oracle port) over `batch` on `nthreads` host threads; seconds."""
    import checkers as ck
    from bsalign_b200 import synth
    kind = "reference" if ck.have_ref() else "port"
    fn = ck.ref_batch if kind == "reference" else ck.oracle_batch
    mtx = synth.score_matrix(*MATRIX)
    res, cigs, _ = fn(w["kind"], batch, w["mode"], w["bandwidth"], mtx, GAPS, nthreads=nthreads, repeat=repeat, want_cigar=True)
    if want_cigars:
        return ck.last_call_seconds, kind, res, cigs
    return ck.last_call_seconds, kind, res  # the C call only: results + cigars written to caller arenas


def poa_make_batch(w, seed, njobs):
    """njobs sweep jobs: `distinct` different synthetic graphs (bsalign_b200/synth_poa.py), tiled; every job has its own arenas."""
    from bsalign_b200 import poa, synth_poa
    d = min(w["distinct"], njobs)
    protos = [synth_poa.make_sweep_job(seed * 1000 + i, tlen=w["qlen"], ngraph=21, par=dict(bandwidth=w["bandwidth"], alnmode=w["mode"]),
                                       p_sub=w["err"][0], p_ins=w["err"][1], p_del=w["err"][2]) for i in range(d)]
    return poa.SweepBatch([protos[i % d] for i in range(njobs)]), protos


def poa_cpu_run(batch, nthreads):
    """The oracle port of align_rd_bspoacore (oracle/bsalign_oracle.c:bso_poa_sweep_batch) on nthreads host threads."""
    import poa_jobs as pj
    r = pj.oracle_sweep_batch(batch, nthreads=nthreads, want_rows=False)
    return r["seconds"], r


def poa_reference_run(w, njobs, nthreads, seed=4000):
    """The UNMODIFIED reference (oracle/_ref: bsref_poa_time): njobs whole BSPOA jobs (64 reads x qlen, DEFAULT_BSPOA_PAR, realn = 0), one per
    host thread at a time; only the CPU time inside align_rd_bspoacore counts.  Returns (GCUPS with all threads busy, description) or None."""
    import checkers as ck
    import poa_jobs as pj
    if not ck.have_ref():
        return None
    sets = [pj.make_reads(64, w["qlen"], seed + i, *w["err"]) for i in range(njobs)]
    r = pj.ref_time_jobs(sets, nthreads)
    val = r["cells"] / (r["dp_seconds"] / min(nthreads, njobs)) / 1e9
    desc = ("%d whole BSPOA jobs (64 reads x %d bp, DEFAULT_BSPOA_PAR) through the unmodified reference on %d host threads; only the thread CPU time inside "
            "align_rd_bspoacore counts: %.2f s for %d row updates + %d merges (%.2f us per update); GCUPS = cells / (that time / threads)"
            % (njobs, w["qlen"], nthreads, r["dp_seconds"], r["nupd"], r["nmrg"], r["dp_seconds"] / max(1, r["nupd"]) * 1e6))
    return val, desc, r


def main_poa(args, w, njobs, ncores, rank, local_rank, world):
    from bsalign_b200 import poa
    bw = w["bandwidth"]
    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's sweep lives inside BSPOA objects (graph surgery, consensus: out of scope); its CPU arm is the oracle port of
        # align_rd_bspoacore, pinned against the reference's own sweep dumps (tests/test_poa.py)
        sample = args.cpu_sample or ncores
        vals, kind, dt = [], "reference", 0.0
        for _ in range(max(1, args.steps)):
            rr = poa_reference_run(w, sample, ncores)
            if rr is None:
                break
            vals.append(rr[0]); dt += rr[2]["wall"] / max(1, args.steps)
        if vals:
            val = float(np.mean(vals))
            sample_desc = rr[1]
        else:   # the compiled reference did not travel: the oracle port of the sweep on synthetic graphs
            kind = "port"
            sample = args.cpu_sample or 2 * ncores
            batch, _ = poa_make_batch(w, 1, sample)
            for _ in range(args.steps):
                d, r = poa_cpu_run(batch, ncores)
                dt += d / args.steps
            val = int(r["ops"][:, 0].sum()) * bw / dt / 1e9
            sample_desc = "%d synthetic sweep jobs per step through the scalar oracle port, %d host threads" % (sample, ncores)
        print(json.dumps({
            "impl": "reference", "metric": "GCUPS", "value": val, "unit": "GCUPS (1e9 band cells/s)", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8", "data": "synthetic",
            "config": {"workload": "c5: %s" % w["desc"], "jobs_per_step": sample},
            "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": ncores, "kind": kind, "sample": sample_desc},
            "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return
    import torch
    import torch.distributed as dist
    from bsalign_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    batch, protos = poa_make_batch(w, 1 + rank, njobs)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    for name in ("par", "queries", "qoff", "slen", "node_off", "base", "bonus", "rpos", "nct", "eoff", "edge_off", "edst", "head", "tail",
                 "reoff", "redge_off", "resrc", "recov"):
        setattr(batch, name, pin(getattr(batch, name)))
    ctx = api.Context(local_rank)
    # ---- kernel-only leg: jobs resident in HBM -------------------------------------------------------
    rs = poa.ResidentSweeps(ctx, batch)
    rs.attach_reverse()          # the run = sweep + the walk of alignment2graph_bspoa on the device
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        rs.run()
    barrier()
    sampler.mark()
    dev_ms = sweep_ms = walk_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rs.run()
        tm = ctx.timing()
        dev_ms += tm["run_ms"]; sweep_ms += tm["forward_ms"]; walk_ms += tm["traceback_ms"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    res = rs.fetch(want_rows=False)
    rs.free()
    nupd, nmrg = int(res.ops[:, 0].sum()), int(res.ops[:, 1].sum())
    cells = nupd * bw
    blk = int(batch.blk[0])
    alg_bytes = (2 * nupd + 3 * nmrg) * blk                    # update reads + writes a block, merge touches three
    step_ms = reduce(dev_ms / args.steps, dist.ReduceOp.MAX if world > 1 else None)
    total_cells = reduce(cells, dist.ReduceOp.SUM if world > 1 else None)
    value = total_cells / (step_ms * 1e-3) / 1e9
    # ---- e2e leg: host buffers through the one-shot C-ABI call: graphs up, sweep + walk, per-read-position matches + best ends back
    # (what include/bsalign_b200_poa_compat.h calls per round).  e2e_rows = the older boundary that returns every row block for the host
    # traceback of the reference (15 MB per job).
    match = pin(np.zeros(int(batch.slen.sum()), dtype=np.int32))
    for _ in range(min(args.warmup, 2)):
        r = poa.poa_align_batch(ctx, batch, match=match)
    barrier()
    t0 = time.perf_counter()
    parts = {"h2d_ms": 0.0, "run_ms": 0.0, "d2h_ms": 0.0}
    for _ in range(args.steps):
        r = poa.poa_align_batch(ctx, batch, match=match)
        tm2 = ctx.timing()
        for k in parts:
            parts[k] += tm2[k] / args.steps
    barrier()
    e2e_ms = reduce((time.perf_counter() - t0) / args.steps * 1e3, dist.ReduceOp.MAX if world > 1 else None)
    e2e_val = total_cells / (e2e_ms * 1e-3) / 1e9
    assert np.array_equal(r.best, res.best), "e2e and resident runs disagree"
    assert np.array_equal(r.trace, res.trace) and int((r.trace[:, 7] != 0).sum()) == 0, "device walk: e2e and resident runs disagree"
    rows = pin(np.zeros(int(batch.row_off[-1]), dtype=np.uint8))
    rr = poa.poa_rows_batch(ctx, batch, rows=rows)
    barrier()
    t0 = time.perf_counter()
    rr = poa.poa_rows_batch(ctx, batch, rows=rows)
    barrier()
    e2e_rows_ms = (time.perf_counter() - t0) * 1e3
    tm3 = ctx.timing()
    r.rows = rr.rows
    # ---- parity: distinct prototypes against the oracle (best ends + every row block) --------------------
    checked = None
    if args.check:
        import poa_jobs as pj
        k = min(args.check, len(protos))
        ob = pj.oracle_sweep_batch(poa.SweepBatch(protos[:k]), nthreads=ncores)
        ok = True
        for i in range(k):
            lin, ub = r.linear(i)
            m = ob["done"][int(batch.node_off[i]):int(batch.node_off[i + 1])].astype(bool)
            ok = ok and np.array_equal(r.best[i], ob["best"][i]) and np.array_equal(lin[m], ob["rows"][i][m]) \
                and np.array_equal(ub[m], ob["ub"][int(batch.node_off[i]):int(batch.node_off[i + 1])][m])
            import test_gpu_poa as tg          # the walk against the oracle's restatement of alignment2graph_bspoa's decisions
            om, oo = pj.oracle_backtrace(tg._as_dump_like(protos[i]), ob["rows"][i], ob["ub"][int(batch.node_off[i]):int(batch.node_off[i + 1])],
                                         int(r.best[i][1]), int(r.best[i][2]))
            ok = ok and np.array_equal(om, r.match(i)) and np.array_equal(oo, r.trace[i])
        checked = {"jobs": k, "bit_exact": bool(ok)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes * args.steps / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0
    traffic, traffic_src = None, None
    try:   # DRAM bytes per launch: ncu-measured ratio (profiles/traffic_r1.json) x this launch's algorithmic bytes
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r1.json")))["poa_sweep_fast_kernel"]
        traffic = tr["ratio"] * alg_bytes
        traffic_src = "ncu dram__bytes_read+write / algorithmic = %.3f (%s), scaled to this launch" % (tr["ratio"], tr["capture"])
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)", "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "poa_sweep_fast_kernel<2,1> (register-resident sweep; jobs with other bands would run poa_sweep_kernel)", "launches_per_step": 1,
                "algorithmic_bytes_per_launch": int(alg_bytes),
                "kernel_ms_per_launch": sweep_ms / args.steps, "row_updates": nupd, "row_merges": nmrg, "block_bytes": blk,
                "walk_kernel_ms_per_step": walk_ms / args.steps}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rr = poa_reference_run(w, args.cpu_sample or ncores, ncores)
        if rr is not None:
            cpu = {"value": rr[0], "unit": "GCUPS", "cores": ncores, "kind": "reference", "sample": rr[1]}
        else:
            sample = min(njobs, args.cpu_sample or 2 * ncores)
            sub = poa.SweepBatch([protos[i % len(protos)] for i in range(sample)])
            dt, cr = poa_cpu_run(sub, ncores)
            cpu = {"value": int(cr["ops"][:, 0].sum()) * bw / dt / 1e9, "unit": "GCUPS", "cores": ncores, "kind": "port",
                   "sample": "first %d sweep jobs of the same batch through the scalar oracle port, %d host threads, %.1f s" % (sample, ncores, dt),
                   "results_equal_gpu": bool(np.array_equal(cr["best"], r.best[:sample]))}
    if rank == 0:
        print(json.dumps({
            "metric": "GCUPS", "value": value, "unit": "GCUPS (1e9 band cells/s)", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": "c5: %s" % w["desc"], "jobs_per_gpu": njobs, "distinct_graphs": len(protos), "nodes_per_job": int(batch.node_off[1]),
                       "l2": "graphs (%.0f MB) and the %.1f GB of row blocks written per step exceed the 126 MB L2" % (
                           (batch.rpos.nbytes * 3 + batch.edst.nbytes) / 1e6, float(batch.row_off[-1]) / 1e9),
                       "timing": "CUDA events on the library stream; max over ranks", "wall_ms_per_step": wall / args.steps * 1e3,
                       "ms_of_dp_per_64_read_job_at_this_batch": step_ms * 39},
            "e2e": {"value": e2e_val, "unit": "GCUPS", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(tm2["h2d_bytes"]), "d2h_bytes_per_step": int(tm2["d2h_bytes"]),
                    "device_parts_ms": parts, "returns": "per-read-position matched node + counts (device-side walk of alignment2graph_bspoa)",
                    "e2e_rows": {"value": total_cells / world / (e2e_rows_ms * 1e-3) / 1e9 * world, "ms_per_step": e2e_rows_ms, "d2h_bytes_per_step": int(tm3["d2h_bytes"]),
                                 "returns": "every node row block (host traceback of the reference)"}},
            "gpu_launches": 4 * args.steps, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "parity": {"nonzero_status_jobs": int((r.status != 0).sum()), "score_checksum": int(r.best[:, 0].astype(np.int64).sum()), "checked": checked},
        }))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_kmer_edit(ctx, ncores, pairs=1000000, qlen=300, ksz=13, steps=2, warmup=2, cpu_seconds=5.0, check=256):
    """the k-mer guided edit (kmer_striped_seqedit_pairwise, bsalign.h:1209; `bsalign edit -m kmer -k 13`) on the
    configs[3] batch shape.  No band-cell count describes this path (the anchors replace most of the DP), so the unit is pairs per second:
    kernel-only from the CUDA events of the call, end to end = wall time of ONE C-ABI call with pinned host buffers; the reference is the
    unmodified kmer_striped_seqedit_pairwise on all host cores."""
    import torch
    import checkers as ck
    from bsalign_b200 import api, synth
    w = WORKLOADS["c4"]
    batch = synth.make_pairs(pairs, qlen, 1000, *w["err"])
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    hb = synth.PairBatch.__new__(synth.PairBatch)
    hb.seqs, hb.qoff, hb.qlen, hb.toff, hb.tlen = pin(batch.seqs), pin(batch.qoff), pin(batch.qlen), pin(batch.toff), pin(batch.tlen)
    outbuf = tuple(pin(a) if a is not None else None for a in api._alloc_out(hb, True))
    got = None
    walls, kern, fb = [], [], []
    # kernel-only: the unpipelined call (one launch over the whole batch, CUDA events around it); end to end: the call as a user makes it
    # (batches of this size are pipelined in chunks over two streams inside the call)
    for pipe in (False, True):
        if pipe:
            os.environ.pop("BSB200_NOPIPE", None)
        else:
            os.environ["BSB200_NOPIPE"] = "1"
        for it in range(warmup + steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            got = ctx.kmer_edit_batch(hb, ksz, dense=True, out=outbuf)
            dt = time.perf_counter() - t0
            tm = ctx.timing()
            if it >= warmup:
                if pipe:
                    walls.append(dt); fb.append(tm["waves"])
                else:
                    kern.append(tm["forward_ms"] + tm["traceback_ms"])
    wall = sum(walls) / len(walls); kms = sum(kern) / len(kern)
    out = {"workload": "c4k: configs[3] batch shape (1M pairs 300bp x 300bp) through the k-mer guided edit, k = %d" % ksz, "pairs": int(batch.n),
           "unit": "Mpairs/s", "value": batch.n / (kms * 1e-3) / 1e6, "kernel_ms_per_step": kms, "e2e": batch.n / wall / 1e6, "e2e_ms_per_step": wall * 1e3,
           "pairs_without_anchors": int(fb[-1]), "gpu_launches": int(tm["forward_launches"] + tm["other_launches"]),
           "h2d_bytes_per_step": int(tm["h2d_bytes"]), "d2h_bytes_per_step": int(tm["d2h_bytes"]), "steps": steps, "warmup": warmup}
    gc = got.cigars()
    m = min(check, batch.n)
    sub = synth.PairBatch(batch.seqs, batch.qoff[:m], batch.qlen[:m], batch.toff[:m], batch.tlen[:m])
    exp, ecg, _ = ck.kmer_batch("oracle", sub, ksz, nthreads=ncores)
    ok = bool(np.array_equal(exp, got.results[:m]) and all(np.array_equal(ecg[i], gc[i]) for i in range(m)))
    out["parity"] = {"checked": {"pairs": m, "against": "oracle", "bit_exact": ok}}
    if cpu_seconds and ck.have_ref():
        m = min(batch.n, 20000)
        sub = synth.PairBatch(batch.seqs, batch.qoff[:m], batch.qlen[:m], batch.toff[:m], batch.tlen[:m])
        t0 = time.perf_counter()
        r, c, _ = ck.kmer_batch("ref", sub, ksz, nthreads=ncores)
        dt = time.perf_counter() - t0
        rate = m / dt / 1e6
        out["cpu"] = {"value": rate, "unit": "Mpairs/s", "cores": ncores, "kind": "reference", "sample": "first %d pairs, %d host threads, %.1f s" % (m, ncores, dt),
                      "results_equal_gpu": bool(np.array_equal(r, got.results[:m]) and all(np.array_equal(c[i], gc[i]) for i in range(m)))}
        out["speedup_vs_cpu"] = out["value"] / rate
        out["e2e_speedup_vs_cpu"] = out["e2e"] / rate
    return out


def run_remsa(ctx, ncores, njobs=1000, mlen=20000, bw=32, distinct=16, steps=2, warmup=2, cpu_seconds=5.0, check=4):
    """SURVEY section 8 row f2 (first part): the DP + walk of remsa_pedit_rd_bspoacore (bspoa.h:3916-4045) for one re-alignment round of
    `njobs` lock-step BSPOA objects (one read each; 15 kb reads give an MSA of ~20k columns; band editbw / 2 = 32 cells).  Synthetic inputs
    in the reference's layout (bsalign_b200/synth_remsa.py; `distinct` different jobs, repeated).  Unit: GCUPS over the band cells of the
    anti-diagonals.  CPU: the oracle's scalar port on all host cores (kind "port": the reference's SSE kernel cannot be called on its own)."""
    import torch
    import remsa_jobs as rj
    from bsalign_b200 import api, synth_remsa
    base = [synth_remsa.make_job(mlen, bw=bw, seed=7000 + k) for k in range(distinct)]
    jobs = [base[k % distinct] for k in range(njobs)]
    rb = api.RemsaBatch(jobs)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    rb.arena = pin(rb.arena)
    outbuf = (pin(np.zeros(max(rb.match_ints, 1), dtype=np.int32)), pin(np.zeros((njobs, 4), dtype=np.int32)))
    walls, kern = [], []
    for it in range(warmup + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ms, res, _ = api.remsa_batch(ctx, rb, out=outbuf)
        dt = time.perf_counter() - t0
        tm = ctx.timing()
        if it >= warmup:
            walls.append(dt); kern.append(tm["forward_ms"])
    cells = float(tm["cells"])
    kms, wall = sum(kern) / len(kern), sum(walls) / len(walls)
    out = {"workload": "c5r: one re-alignment round (remsa_pedit_rd_bspoacore) over %d lock-step MSA jobs, MSA of %d columns, band %d" % (njobs, mlen, bw),
           "jobs": njobs, "unit": "GCUPS", "value": cells / (kms * 1e-3) / 1e9, "kernel_ms_per_step": kms, "e2e": cells / wall / 1e9, "e2e_ms_per_step": wall * 1e3,
           "gpu_launches": 1, "h2d_bytes_per_step": int(tm["h2d_bytes"]), "d2h_bytes_per_step": int(tm["d2h_bytes"]), "matrix_bytes_per_step": int(tm["trace_bytes"]),
           "steps": steps, "warmup": warmup}
    ok = True
    for k in range(min(check, distinct)):
        _, _, omatch, oscr, oerr = rj.oracle_core(base[k])
        ok = ok and oerr == 0 and np.array_equal(omatch, ms[k]) and int(res[k, 0]) == oscr and int(res[k, 1]) == 0
    out["parity"] = {"checked": {"jobs": min(check, distinct), "against": "oracle (pinned to the instrumented reference)", "bit_exact": bool(ok)}}
    if cpu_seconds:
        import concurrent.futures as cf
        m = 8 * max(ncores, 16)
        t0 = time.perf_counter()
        with cf.ThreadPoolExecutor(ncores) as ex:
            list(ex.map(lambda k: rj.oracle_core(base[k % distinct])[3], range(m)))
        dt = time.perf_counter() - t0
        rate = cells / njobs * m / dt / 1e9
        out["cpu"] = {"value": rate, "unit": "GCUPS", "cores": ncores, "kind": "port", "sample": "%d jobs, %d host threads, %.1f s (the oracle's scalar restatement)" % (m, ncores, dt)}
        out["speedup_vs_cpu"] = out["value"] / rate
    return out


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def profile_facts(kind, waved):
    """Per-kernel facts measured once under ncu and committed under profiles/ (traffic ratio, ALU-pipe share): profiles/kernel_facts_r2.json."""
    try:
        facts = json.load(open(os.path.join(ROOT, "profiles", "kernel_facts_r2.json")))
        return facts["epi8_wave_kernel" if (kind == "epi8" and waved) else ("epi8_forward_kernel" if kind == "epi8" else "edit_kernel")]
    except Exception:
        return None


def config_of(name, w, pairs):
    """The same keys in both arms (the driver compares them)."""
    return {"workload": "%s: %s" % (name, w["desc"]), "pairs": int(pairs), "qlen": w["qlen"], "mode": w["mode"], "bandwidth": w["bandwidth"],
            "matrix": list(MATRIX), "gaps": list(GAPS), "seed": 1000}


def compare_with_cpu(r, idx, exp, ecg):
    """Bit-exact check of result records AND cigar words of the pairs idx; returns the number of pairs that differ."""
    idx = np.asarray(idx, dtype=np.int64)
    bad = (r.results[idx] != np.asarray(exp)).any(axis=1)
    bad |= np.asarray(r.ncigar)[idx].astype(np.int64) != np.array([len(c) for c in ecg], dtype=np.int64)
    for k in np.nonzero(~bad)[0]:
        if not np.array_equal(r.cigar(int(idx[k])), ecg[k]):
            bad[k] = True
    return int(bad.sum())


def run_workload(ctx, name, w, pairs, args, ncores, rank, world, local_rank, dist, torch, cpu_seconds, check):
    """One workload on this rank's GPU (N = 1) or sharded over all ranks (N > 1: the batch lives on rank 0).  Returns the JSON fields."""
    from bsalign_b200 import api, synth, shard
    import checkers as ck
    mtx = synth.score_matrix(*MATRIX)
    kind = w["kind"]
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allred(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    batch = make_batch(w, 1000, pairs) if rank == 0 else None       # ONE batch; with N > 1 it lives on rank 0 only
    cells = nominal_cells(w, batch) if rank == 0 else 0
    sampler = ClockSampler(local_rank)
    out = {}
    if world == 1:
        hb = synth.PairBatch.__new__(synth.PairBatch)
        hb.seqs, hb.qoff, hb.qlen, hb.toff, hb.tlen = pin(batch.seqs), pin(batch.qoff), pin(batch.qlen), pin(batch.toff), pin(batch.tlen)
        outbuf = tuple(pin(a) if a is not None else None for a in api._alloc_out(hb, True))
        rb = ctx.upload(kind, hb, w["mode"], w["bandwidth"], mtx, GAPS, want_cigar=True)
    else:
        # ---- scatter once, shards stay resident for the kernel-only leg ------------------------------------------------
        hdr = torch.zeros(1, dtype=torch.int64, device=dev)
        if rank == 0:
            hdr[0] = batch.n
        dist.broadcast(hdr, 0)
        n = int(hdr.item())
        lens = torch.zeros((2, n), dtype=torch.int32, device=dev)
        if rank == 0:
            lens[0] = torch.from_numpy(batch.qlen.astype(np.int32)).to(dev); lens[1] = torch.from_numpy(batch.tlen.astype(np.int32)).to(dev)
        dist.broadcast(lens, 0)
        lh = lens.cpu().numpy()
        plans = shard.plan_shards(lh[0].astype(np.uint32), lh[1].astype(np.uint32), kind, w["bandwidth"], world)
        mine = plans[rank]
        if rank == 0:
            reqs, keep = [], []
            for r in list(range(1, world)) + [0]:
                pb, nb = api.pack_pairs(batch, plans[r]["idx"], nthreads=ncores)
                t = torch.from_numpy(pb.seqs[:max(nb, 1)]).to(dev)
                keep.append(t)
                if r:
                    reqs.append(dist.isend(t, r))
                else:
                    arena = t
            for q in reqs:
                q.wait()
        else:
            arena = torch.empty(max(mine["nbytes"], 1), dtype=torch.uint8, device=dev)
            dist.recv(arena, 0)
        torch.cuda.synchronize()
        rb = ctx.upload_dev(kind, arena.data_ptr(), shard.ShardView(mine), w["mode"], w["bandwidth"], mtx, GAPS, want_cigar=True)
    # ---- kernel-only leg: inputs resident in HBM ------------------------------------------------------------------------
    sampler.start()
    for _ in range(args.warmup):
        rb.run()
    barrier()
    sampler.mark()
    dev_ms = fwd_ms = bt_ms = 0.0
    fwd_launches = bt_launches = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rb.run()
        tm = ctx.timing()
        dev_ms += tm["run_ms"]; fwd_ms += tm["forward_ms"]; bt_ms += tm["traceback_ms"]
        fwd_launches += tm["forward_launches"]; bt_launches += tm["traceback_launches"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    trace_bytes, waves = tm["trace_bytes"], tm["waves"]
    if world == 1:
        last = rb.fetch(out=outbuf)
    rb.free()
    step_ms = allred(dev_ms / args.steps, dist.ReduceOp.MAX if world > 1 else None)
    fwd_step_ms = allred(fwd_ms / args.steps, dist.ReduceOp.MAX if world > 1 else None)
    bt_step_ms = allred(bt_ms / args.steps, dist.ReduceOp.MAX if world > 1 else None)
    total_cells = allred(cells, dist.ReduceOp.SUM if world > 1 else None)
    total_trace = allred(trace_bytes, dist.ReduceOp.SUM if world > 1 else None)
    launches = allred(fwd_launches + bt_launches, dist.ReduceOp.SUM if world > 1 else None)
    value = total_cells / (step_ms * 1e-3) / 1e9

    # ---- e2e leg: host buffers on rank 0 through the public call ---------------------------------------------------------
    e2e_extra = {}
    if world == 1:
        def one_e2e():
            if kind == "epi8":
                return ctx.epi8_batch(hb, w["mode"], w["bandwidth"], mtx, *GAPS, out=outbuf, dense=True)
            return ctx.edit_batch(hb, w["mode"], w["bandwidth"], out=outbuf, dense=True)
        for _ in range(min(args.warmup, 3)):
            one_e2e()
        barrier()
        t0 = time.perf_counter()
        parts = {"h2d_ms": 0.0, "run_ms": 0.0, "d2h_ms": 0.0}
        for _ in range(args.steps):
            r = one_e2e()
            tm2 = ctx.last_timing
            h2d, d2h = tm2["h2d_bytes"], tm2["d2h_bytes"]
            for k in parts:
                parts[k] += tm2[k] / args.steps
        barrier()
        e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
        chunks = int(tm2["waves"]) if kind == "edit" else 1   # large edit batches are pipelined in chunks over two streams inside the call
        e2e_extra = {"device_parts_ms": parts, "host_planning_and_scatter_ms": max(0.0, e2e_ms - sum(parts.values())), "host_buffers": "pinned"}
        if chunks > 1:
            e2e_extra["pipelined_chunks"] = chunks
            e2e_extra["device_parts_note"] = "sums over the chunks; copies, kernels and host planning of different chunks overlap, so the wall time is shorter than the sum"
        # the same call with PAGEABLE caller buffers (what a reference caller holds: plain u1i* arrays)
        pout = api._alloc_out(batch, True)
        for it in range(2):   # one warm-up call (the library allocates its pinned staging buffers on first use), one timed
            t0 = time.perf_counter()
            if kind == "epi8":
                ctx.epi8_batch(batch, w["mode"], w["bandwidth"], mtx, *GAPS, out=pout, dense=True)
            else:
                ctx.edit_batch(batch, w["mode"], w["bandwidth"], out=pout, dense=True)
            torch.cuda.synchronize()
            pg_ms = (time.perf_counter() - t0) * 1e3
        e2e_extra["pageable"] = {"value": total_cells / (pg_ms * 1e-3) / 1e9, "ms_per_step": pg_ms}
        # ... and with the sequences 2-bit packed the way the reference keeps them (BaseBank words, dna.h:63; main.c unpacks a pair per call):
        # one call of bsb200_pairwise_batch_dense_bits, packed words in pinned memory
        bits = pin(api.pack_bits(batch.seqs))
        pk_ms = 0.0
        for it in range(1 + args.steps):
            t0 = time.perf_counter()
            rpk = ctx.dense_bits(kind, bits, hb, w["mode"], w["bandwidth"], mtx, GAPS, out=outbuf)
            tmb = ctx.last_timing
            if it:
                pk_ms += (time.perf_counter() - t0) * 1e3 / args.steps
        assert np.array_equal(rpk.results, last.results), "2-bit packed upload and byte upload disagree"
        e2e_extra["packed_2bit"] = {"value": total_cells / (pk_ms * 1e-3) / 1e9, "ms_per_step": pk_ms, "h2d_bytes_per_step": int(tmb["h2d_bytes"])}
        r = rpk
        assert int(r.results[:, 0].astype(np.int64).sum()) == int(last.results[:, 0].astype(np.int64).sum()), "e2e and resident runs disagree"
    else:
        # rank 0 owns the batch in (pinned) host memory: lengths broadcast, compact arenas scattered over NVLink, every rank aligns its
        # shard in place, records + dense cigars gathered to rank 0 and merged into pair order (bsalign_b200/shard.py)
        if rank == 0:
            hb = synth.PairBatch.__new__(synth.PairBatch)
            hb.seqs, hb.qoff, hb.qlen, hb.toff, hb.tlen = pin(batch.seqs), batch.qoff, batch.qlen, batch.toff, batch.tlen
        else:
            hb = None
        pool = {}

        def pinned(nbytes):   # pinned staging arenas, kept between steps
            k = len(pool_used)
            if k not in pool or pool[k].size < nbytes:
                pool[k] = torch.empty(int(nbytes * 1.05) + 64, dtype=torch.uint8).pin_memory().numpy()
            pool_used.append(k)
            return pool[k]
        al = shard.cuda_aligner(ctx, kind, w["mode"], w["bandwidth"], mtx, GAPS)
        stage = {}
        for it in range(min(args.warmup, 2) + args.steps):
            if it == min(args.warmup, 2):
                barrier()
                t0 = time.perf_counter()
                stage = {}
            pool_used = []
            timers = {}
            res = shard.run_sharded_device(hb, kind, w["bandwidth"], al, dist, device=dev, pinned=pinned, nthreads=ncores, timers=timers, packer=shard.cuda_packer(ctx))
            for k in ("plan", "scatter_issued", "aligned", "gathered", "assembled"):
                if k in timers:
                    stage[k] = stage.get(k, 0.0) + timers[k] / args.steps
        barrier()
        e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
        h2d = int(batch.seqs.nbytes + 8 * batch.n) if rank == 0 else 0
        d2h = 0
        if rank == 0:
            results, status, ncigar, dense, goff = res
            d2h = int(results.nbytes + status.nbytes + ncigar.nbytes + dense.nbytes)
            r = api.BatchResult(results, dense, goff, ncigar, status)
            e2e_extra = {"scatter_bytes_over_nvlink": int(timers.get("scatter_bytes", 0)), "gather_bytes_over_nvlink": int(timers.get("gather_bytes", 0)),
                         "rank0_stage_ms_cumulative": stage,
                         "path": "rank 0 host batch -> broadcast lengths -> pack + H2D + NCCL send of compact shard arenas -> bsb200_batch_upload_dev / run / fetch_dense_dev on every rank -> NCCL send of records + dense cigars -> pair-ordered merge on rank 0"}
    e2e_ms = allred(e2e_ms, dist.ReduceOp.MAX if world > 1 else None)
    e2e_val = total_cells / (e2e_ms * 1e-3) / 1e9

    if rank == 0:
        checksum = int(r.results[:, 0].astype(np.int64).sum())
        bad_status = int((r.status != 0).sum())
        # ---- parity: `check` pairs of the last e2e step against the CPU restatement, result records AND cigar words ----
        checked = None
        if check:
            idx = np.unique(np.linspace(0, batch.n - 1, min(check, batch.n)).astype(np.int64))
            exp, ecg, _ = ck.oracle_batch(kind, batch.subset(idx), w["mode"], w["bandwidth"], mtx, GAPS, nthreads=ncores)
            nbad = compare_with_cpu(r, idx, exp, ecg)
            checked = {"pairs": int(len(idx)), "fields": "10 result ints + every cigar word", "against": "oracle port (pinned to the reference)", "mismatches": int(nbad), "bit_exact": nbad == 0}
        # ---- roofline of the dominant kernel (forward): algorithmic trace bytes / its event time, max over ranks ----------
        peaks = load_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = total_trace / world / (fwd_step_ms * 1e-3) / 1e9 if fwd_step_ms > 0 else 0.0    # per GPU
        waved = kind == "epi8" and fwd_launches >= 2 * args.steps * max(1, waves) and w["bandwidth"] == 0
        kname = ("epi8_wave_kernel" if waved else "epi8_forward_kernel") if kind == "epi8" else "edit_kernel"
        facts = profile_facts(kind, waved)
        traffic = traffic_src = int_issue = None
        if facts:
            traffic = facts["dram_over_algorithmic"] * total_trace / world / max(1, waves)
            traffic_src = "ncu dram__bytes_read+write / algorithmic = %.3f (%s), scaled to this launch" % (facts["dram_over_algorithmic"], facts["capture"])
            # ALU-pipe share measured under ncu, scaled by the ratio of the live kernel time per cell to the profiled one
            live = fwd_step_ms * 1e-3 / (total_cells / world)
            if facts.get("alu_pipe_pct") and facts.get("seconds_per_cell"):
                int_issue = facts["alu_pipe_pct"] / 100.0 * facts["seconds_per_cell"] / live
        out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                           "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                           "traffic": traffic, "traffic_source": traffic_src, "kernel": kname, "launches_per_step": int(waves),
                           "int_issue_frac": int_issue,
                           "int_issue_note": "the kernel's binding bound is integer issue, not HBM: ALU-pipe utilisation (sm__inst_executed_pipe_alu, ncu) of the committed capture scaled to this run's time per cell",
                           "algorithmic_bytes_per_launch": int(total_trace / world / max(1, waves)), "kernel_ms_per_launch": fwd_step_ms / max(1, waves),
                           "algorithmic_bytes_per_step": int(total_trace / world), "kernel_ms_per_step": fwd_step_ms,
                           "traceback_ms_per_step": bt_step_ms, "waves_per_step": int(waves), "per": "GPU (max over ranks)"}
        # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ----------------------------------------------------
        cpu = None
        if world == 1 and cpu_seconds > 0:
            per_core = {"c2": 2.3, "c3": 1.5, "c4": 1.1, "g10k": 2.1}[name]
            sample = args.cpu_sample or int(max(ncores, min(pairs, cpu_seconds * per_core * 1e9 * ncores / (cells / batch.n))))
            sub = batch.subset(np.arange(sample))
            cpu_reference_run(w, sub.subset(np.arange(min(sample, 2 * ncores))), ncores)
            dt, ckind, cres, ccg = cpu_reference_run(w, sub, ncores, want_cigars=True)
            nbad = compare_with_cpu(r, np.arange(sample), cres, ccg)
            cpu = {"value": nominal_cells(w, sub) / dt / 1e9, "unit": "GCUPS", "cores": ncores, "kind": ckind,
                   "sample": "first %d pairs of the same batch, %d host threads, %.1f s" % (sample, ncores, dt),
                   "results_equal_gpu": nbad == 0, "compared": "10 result ints + every cigar word of %d pairs" % sample}
        out.update({
            "value": value, "ms_per_step": step_ms,
            "config": config_of(name, w, pairs),
            "e2e": dict({"value": e2e_val, "unit": "GCUPS", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}, **e2e_extra),
            "gpu_launches": int(launches), "cpu_baseline": cpu, "clocks": clocks,
            "parity": {"nonzero_status_pairs": bad_status, "score_checksum": checksum, "checked": checked},
            "notes": {"l2": "inputs (%.0f MB) and the %.1f GB traceback store written per step both exceed the 126 MB L2" % (batch.seqs.nbytes / 1e6, total_trace / 1e9),
                      "timing": "CUDA events on the library stream; max over ranks", "wall_ms_per_step": wall / args.steps * 1e3,
                      "sharding": None if world == 1 else "one batch on rank 0, %d shards of equal DP cells (bsalign_b200/shard.py); value = kernels on resident shards, e2e = the whole scatter/align/gather call" % world},
        })
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs in the batch (default: the workload's)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU baseline sample (default: sized for ~10-20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short g10k / c3 / c4 runs reported under `secondary`")
    ap.add_argument("--check", type=int, default=1024, help="verify this many pairs of the last step (results + cigars) against the oracle")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = WORKLOADS[args.workload]
    pairs = args.pairs or w["pairs"]

    def g10k_pairs():
        # a 10 kb x 10 kb pair needs 215 MB of traceback store: take what ONE wave of the library's default budget (90 % of free HBM) seats
        try:
            import torch
            free, _total = torch.cuda.mem_get_info(local_rank)
            return max(148, int(0.90 * free * 0.97 / 218.0e6)) * world
        except Exception:
            return WORKLOADS["g10k"]["pairs"]
    if args.workload == "g10k" and not args.pairs and args.impl == "ours":
        pairs = g10k_pairs()
    ncores = os.cpu_count() or 1
    if w["kind"] == "poa":
        return main_poa(args, w, pairs, ncores, rank, local_rank, world)

    # ------------------------------------------------------------------ reference arm (CPU) -----------
    if args.impl == "reference":
        if rank != 0:
            return
        # bounded sample: sized from the survey's single-core rates so one step is a few seconds on all cores
        per_core_gcups = {"c2": 2.3, "c3": 1.5, "c4": 1.1, "g10k": 2.1}[args.workload]
        one = make_batch(w, 999, 8)
        cells_per_pair = nominal_cells(w, one) / one.n
        sample = args.cpu_sample or int(max(ncores, min(pairs, 4.0 * per_core_gcups * 1e9 * ncores / cells_per_pair)))
        batch = make_batch(w, 1000, sample)
        cells = nominal_cells(w, batch)
        for _ in range(args.warmup):
            cpu_reference_run(w, batch.subset(np.arange(min(sample, ncores * 4))), ncores)
        kind, dt = "port", 0.0
        for _ in range(args.steps):
            d, kind, _ = cpu_reference_run(w, batch, ncores)
            dt += d / args.steps
        val = cells / dt / 1e9
        sample_desc = "%d pairs of the %s shape per step, %d host threads" % (sample, args.workload, ncores)
        print(json.dumps({
            "impl": "reference", "metric": "GCUPS", "value": val, "unit": "GCUPS (1e9 band cells/s)", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
            "dtype": "int8" if w["kind"] == "epi8" else "u64 bit-planes", "data": "synthetic",
            "config": config_of(args.workload, w, pairs),
            "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": ncores, "kind": kind, "sample": sample_desc},
            "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ our arm (GPU) -----------------
    import torch
    import torch.distributed as dist
    from bsalign_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator comes up; stdout must carry the JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    ctx = api.Context(local_rank)
    main_out = run_workload(ctx, args.workload, w, pairs, args, ncores, rank, world, local_rank, dist, torch,
                            0 if args.no_cpu_baseline else 12.0, args.check)
    # ---- secondary: short runs of the other pairwise workloads in the same process (N = 1 only), so that their numbers are
    # driver-visible next to the headline: g10k (north_star target: >= 10x the reference's CPU GCUPS), c3, c4 ---------------------
    secondary = None
    if world == 1 and not args.no_secondary and args.workload == "c2" and not args.pairs:
        secondary = {}
        sargs = argparse.Namespace(**vars(args))
        sargs.steps, sargs.warmup, sargs.cpu_sample = 2, 2, 0
        for name in ("g10k", "c3", "c4"):
            sw = WORKLOADS[name]
            try:
                ctx.trim()   # hand the previous workload's traceback arena back, so that the next one plans against the whole device
                sp = g10k_pairs() if name == "g10k" else sw["pairs"]
                so = run_workload(ctx, name, sw, sp, sargs, ncores, rank, world, local_rank, dist, torch, 0 if args.no_cpu_baseline else 5.0, min(args.check, 256))
                cpu = so.get("cpu_baseline")
                secondary[name] = {"workload": so["config"]["workload"], "pairs": so["config"]["pairs"], "value": so["value"], "ms_per_step": so["ms_per_step"],
                                   "e2e": so["e2e"]["value"], "e2e_ms_per_step": so["e2e"]["ms_per_step"], "e2e_pageable": so["e2e"].get("pageable", {}).get("value"), "e2e_packed_2bit": so["e2e"].get("packed_2bit", {}).get("value"),
                                   "roofline_frac": so["roofline"]["frac"], "kernel": so["roofline"]["kernel"], "kernel_ms_per_step": so["roofline"]["kernel_ms_per_step"],
                                   "traceback_ms_per_step": so["roofline"]["traceback_ms_per_step"],
                                   "cpu": None if not cpu else {"value": cpu["value"], "cores": cpu["cores"], "kind": cpu["kind"], "sample": cpu["sample"], "results_equal_gpu": cpu["results_equal_gpu"]},
                                   "speedup_vs_cpu": None if not cpu else so["value"] / cpu["value"], "e2e_speedup_vs_cpu": None if not cpu else so["e2e"]["value"] / cpu["value"],
                                   "parity": so["parity"], "steps": sargs.steps, "warmup": sargs.warmup}
            except Exception as e:   # a secondary run must never take the headline line with it
                secondary[name] = {"error": repr(e)}
        try:
            ctx.trim()
            secondary["c4k"] = run_kmer_edit(ctx, ncores, cpu_seconds=0 if args.no_cpu_baseline else 5.0, check=min(args.check, 256))
        except Exception as e:
            secondary["c4k"] = {"error": repr(e)}
        try:
            ctx.trim()
            secondary["c5r"] = run_remsa(ctx, ncores, cpu_seconds=0 if args.no_cpu_baseline else 5.0)
        except Exception as e:
            secondary["c5r"] = {"error": repr(e)}
    if rank == 0:
        line = {"metric": "GCUPS", "value": main_out["value"], "unit": "GCUPS (1e9 band cells/s)", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": main_out["ms_per_step"], "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
                "dtype": "int8" if w["kind"] == "epi8" else "u64 bit-planes", "data": "synthetic"}
        for k in ("config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "parity", "notes"):
            line[k] = main_out[k]
        line["secondary"] = secondary
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
