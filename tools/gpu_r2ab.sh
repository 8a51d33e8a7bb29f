#!/bin/bash
# round-2 evidence run on one B200: GPU tests, default bench line, launch list, memory-bound rooflines (+ DRAM traffic), C4 pass
mkdir -p gpurun_out
T=r2ab
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.txt 2>&1; tail -3 gpurun_out/${T}_pytest.txt
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
echo "bench rc $?"; head -c 600 gpurun_out/${T}_bench_default.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_c2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-cycle --no-extra --no-parity > gpurun_out/${T}_ncu_launch.log 2>&1
echo "ncu launch rc $?"
timeout 400 python tools/membound_roofline.py > gpurun_out/${T}_membound_roofline.json 2> gpurun_out/${T}_membound.err
echo "membound rc $?"; tail -3 gpurun_out/${T}_membound.err
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/${T}_membound_ncu.csv python tools/membound_roofline.py --small --reps 1 > gpurun_out/${T}_membound_ncu.log 2>&1
echo "membound ncu rc $?"
timeout 400 python bench.py --workload c4 --subsample 2 --steps 1 --warmup 1 --no-cpu --no-e2e --no-cycle \
  > gpurun_out/${T}_bench_c4_half.json 2> gpurun_out/${T}_bench_c4_half.err
echo "c4 rc $?"; head -c 400 gpurun_out/${T}_bench_c4_half.json; echo
