#!/bin/bash
mkdir -p gpurun_out
T=r2ai
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_additive_inflation.py tests/test_obs_qc.py -m gpu -q -x > gpurun_out/${T}_pytest.txt 2>&1; tail -3 gpurun_out/${T}_pytest.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-cycle --no-extra > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ai_bench_c2.json").read().strip().splitlines()[-1])
print("c2 ms", d["ms_per_step"], "kernel", d["kernel_ms_per_step"], "frac", d["roofline"]["frac"], d["parity"])
PY
timeout 200 python tools/membound_roofline.py --reps 3 > gpurun_out/${T}_membound.json 2> gpurun_out/${T}_membound.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2ai_membound.json"))
print([(r["kernel"][:26], r.get("ms"), r.get("frac")) for r in d["rows"][:6]])
PY
