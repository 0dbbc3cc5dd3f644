#!/bin/bash
# A/B of the two-pass kernel's row store: LDS.128 -> STG.128 loop (BSB200_BULK_STORE=0) against cp.async.bulk copies out of the
# shared-memory images (=1).  Parity first (the GPU parity tests with the bulk store on), then alternating bench runs of
# config 3 (moving band 512: W = 32, 512-byte images) and of config 2 forced onto the two-pass kernel (W = 63, 1 KB images).
out=gpurun_out/ab_bulk_store; mkdir -p $out
BSB200_BULK_STORE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q > $out/pytest_bulk1.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_bulk1.log
for rep in 1 2; do for v in 0 1; do
  BSB200_BULK_STORE=$v timeout 200 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --check 256 > $out/c3_bulk${v}_rep$rep.json 2> $out/c3_bulk${v}_rep$rep.err
  BSB200_NOWAVE=1 BSB200_BULK_STORE=$v timeout 200 python bench.py --workload c2 --pairs 20000 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --check 256 > $out/c2twopass_bulk${v}_rep$rep.json 2> $out/c2twopass_bulk${v}_rep$rep.err
done; done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_bulk_store/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value', round(d['value'],1), 'kernel_ms', round(d['roofline']['kernel_ms_per_step'],3), 'tb_ms', round(d['roofline']['traceback_ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'parity', d['parity']['checked']['bit_exact'], d['roofline']['kernel'])
    except Exception as e: print(f, 'ERR', e)
P
tail -3 $out/pytest_bulk1.log
