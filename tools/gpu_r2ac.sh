#!/bin/bash
# NOBS_OUT / additive inflation GPU tests; ncu launch list of the bench restricted to the repo's own kernels; DRAM traffic of
# the memory-bound kernels; full capture of the one-pass transpose kernel
mkdir -p gpurun_out
T=r2ac
OWN='regex:^(das_|core_kernel|presearch|search_kernel|nobs_out|addi|bucket_|buf_to|clamp_min|dfma|dmma|ensmean|enssprd|exclusive_scan|fill_kernel|grd_|monit_|obs_|obsope|state_trans|tl_|tlc_)'
timeout 300 python -m pytest tests/test_nobs_out.py tests/test_additive_inflation.py -m gpu -x -q > gpurun_out/${T}_pytest.txt 2>&1; tail -5 gpurun_out/${T}_pytest.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OWN" -c 400 --csv --log-file gpurun_out/${T}_launches_bench_c2.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-cycle --no-extra --no-parity > gpurun_out/${T}_ncu_launch.log 2>&1
echo "ncu launch rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$OWN" -c 200 --csv \
  --log-file gpurun_out/${T}_membound_ncu.csv python tools/membound_roofline.py --reps 1 > gpurun_out/${T}_membound_ncu.log 2>&1
echo "membound ncu rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:grd_ens_p2p_kernel -s 8 -c 1 -f -o gpurun_out/r02_scatter_p2p \
  python tools/membound_roofline.py --reps 1 > gpurun_out/${T}_ncu_scatter.log 2>&1
echo "scatter ncu rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:grd_ens_p2p_kernel -s 24 -c 1 -f -o gpurun_out/r02_gather_p2p \
  python tools/membound_roofline.py --reps 1 > gpurun_out/${T}_ncu_gather.log 2>&1
echo "gather ncu rc $?"
ls -la gpurun_out/*.ncu-rep
