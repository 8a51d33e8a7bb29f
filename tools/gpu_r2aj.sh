#!/bin/bash
mkdir -p gpurun_out
T=r2aj
LETKF_B200_ADDINFL_REG=1 timeout 200 python -m pytest tests/test_additive_inflation.py -m gpu -q -x > gpurun_out/${T}_pytest_reg.txt 2>&1; tail -2 gpurun_out/${T}_pytest_reg.txt
LETKF_B200_ADDINFL_REG=1 timeout 200 python tools/membound_roofline.py --reps 3 > gpurun_out/${T}_membound_reg.json 2> gpurun_out/${T}_membound_reg.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2aj_membound_reg.json"))
print("reg path:", [(r["kernel"][:26], r.get("ms"), r.get("frac")) for r in d["rows"] if "additive" in r["kernel"]])
PY
