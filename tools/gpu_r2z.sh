#!/bin/bash
# stagger experiment + ncu --set full captures of the two solver instantiations
mkdir -p gpurun_out
B="--steps 3 --warmup 2 --no-cpu --no-e2e --no-cycle --no-extra --no-parity"
for us in 0 12 25 50; do
  LETKF_B200_STAGGER_US=$us timeout 300 python bench.py --workload c2 $B > gpurun_out/r2z_stag$us.json 2> gpurun_out/r2z_stag$us.err
  python - $us <<'PY'
import json, sys
us = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2z_stag{us}.json").read().strip().splitlines()[-1])
    print("stagger", us, "us: ms %.2f" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"], d["phase_share_rank0"])
except Exception as e:
    print("stagger", us, "FAILED", repr(e))
PY
done
N="--steps 1 --warmup 2 --no-cpu --no-e2e --no-cycle --no-extra --no-parity"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:das_ns_kernel -s 28 -c 1 -f -o gpurun_out/r02_das_ns7 \
  python bench.py --workload c2 --subsample 8 $N > gpurun_out/r2z_ncu7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:das_ns_kernel -s 10 -c 1 -f -o gpurun_out/r02_das_ns13 \
  python bench.py --workload c3small $N > gpurun_out/r2z_ncu13.log 2>&1
ls -la gpurun_out/*.ncu-rep
