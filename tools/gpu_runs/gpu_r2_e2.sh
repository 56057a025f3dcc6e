#!/bin/bash
O=gpurun_out/r2e_n2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_cfg4_parity.py -x -q > $O/pytest_cfg4.log 2>&1; echo "rc=$?" >> $O/pytest_cfg4.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_cfg3.csv python bench.py --workload cfg3_allpairs --steps 8 --warmup 3 --no-cpu > $O/launches_cfg3.log 2>&1
bash tools/gpu_r2_e.sh 2 "2]"
