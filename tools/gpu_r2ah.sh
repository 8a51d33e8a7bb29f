#!/bin/bash
# final evidence run of round 2: full GPU test suite, DRAM traffic of one whole das_letkf step (all solver + pre-search launches), default bench line
mkdir -p gpurun_out
T=r2ah
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.txt 2>&1; tail -3 gpurun_out/${T}_pytest.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:^(das_ns_kernel|presearch_kernel)' --csv --log-file gpurun_out/${T}_das_step_dram_traffic.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-cycle --no-extra --no-parity > gpurun_out/${T}_ncu_traffic.log 2>&1
echo "traffic rc $?"
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
echo "bench rc $?"; head -c 400 gpurun_out/${T}_bench_default.json; echo
