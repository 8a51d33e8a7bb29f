#!/bin/bash
# A/B of kernel variants on one box: tools/gpu_ab.sh <tag> <lib1> [<lib2> ...]   (libs relative to scale_letkf_b200/)
# Every variant runs the C2 bench (3 steps) and the C3-shaped small bench; outputs land in gpurun_out/<tag>_<lib>_*.json
tag=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  name=$(basename "$lib" .so)
  export LETKF_B200_LIB=$PWD/scale_letkf_b200/$lib
  for wl in ${AB_WORKLOADS:-c2 c3small}; do
    timeout 600 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu --no-e2e --no-cycle --no-extra \
      > gpurun_out/${tag}_${name}_${wl}.json 2> gpurun_out/${tag}_${name}_${wl}.err
    python - "$tag" "$name" "$wl" <<'PY'
import json, sys
tag, name, wl = sys.argv[1:4]
try:
    d = json.loads(open(f"gpurun_out/{tag}_{name}_{wl}.json").read().strip().splitlines()[-1])
    print(name, wl, "ms %.2f" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"], d["phase_share_rank0"], d.get("parity"))
except Exception as e:
    print(name, wl, "FAILED", repr(e))
PY
  done
done
