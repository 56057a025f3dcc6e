#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file $O/launches_cfg3.csv python bench.py --workload cfg3_allpairs --steps 8 --warmup 3 --no-cpu > $O/launches_cfg3.log 2>&1
timeout 300 python bench.py --workload cfg3_allpairs --steps 400 --warmup 10 --no-cpu > $O/cfg3_allpairs.json 2> $O/cfg3_allpairs.err
timeout 300 python bench.py --workload cfg2 --steps 20000 --warmup 2000 --no-cpu > $O/cfg2.json 2> $O/cfg2.err
ls -la $O
