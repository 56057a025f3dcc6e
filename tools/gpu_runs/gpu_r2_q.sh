#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_nlist.py tests/test_zz_gpu_ab_switches.py tests/test_abi.py -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary --cpu-budget 3 > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --workload cfg3_nlist --steps 200 --warmup 20 --no-cpu > $O/cfg3_nlist.json 2> $O/cfg3_nlist.err
ls $O
