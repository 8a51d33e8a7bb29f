#!/bin/bash
mkdir -p gpurun_out
T=r2af
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_nobs_out.py tests/test_state_trans.py -m gpu -x -q -k "p2p or nobs or trans" > gpurun_out/${T}_pytest.txt 2>&1; tail -4 gpurun_out/${T}_pytest.txt
timeout 300 python tools/membound_roofline.py --reps 5 > gpurun_out/${T}_membound_tile16.json 2> gpurun_out/${T}_membound_tile16.err; echo "rc $?"
LETKF_B200_P2P_TILE=32 timeout 300 python tools/membound_roofline.py --reps 5 > gpurun_out/${T}_membound_tile32.json 2> gpurun_out/${T}_membound_tile32.err; echo "rc $?"
python - <<'PY'
import json
for v in ("tile16", "tile32"):
    try:
        d = json.load(open(f"gpurun_out/r2af_membound_{v}.json"))
        print(v, [(r["kernel"][:24], r.get("ms"), r.get("frac")) for r in d["rows"] if "p2p" in r["kernel"]])
    except Exception as e:
        print(v, "FAILED", repr(e))
PY
