#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sorted_allpairs.py tests/test_gpu_parity.py tests/test_zz_gpu_ab_switches.py -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
LJMD_SORT_INTERVAL=50 timeout 300 python bench.py --workload cfg3_allpairs --steps 400 --warmup 10 --no-cpu > $O/cfg3_allpairs_resort50.json 2> $O/cfg3_allpairs_resort50.err
timeout 300 python bench.py --workload cfg3_allpairs --steps 400 --warmup 10 --no-cpu > $O/cfg3_allpairs.json 2> $O/cfg3_allpairs.err
ls $O
