#!/bin/bash
# 1-GPU validation of the new defaults (publish in kernel, page-locked caller arrays) + the default bench line
O=gpurun_out/r2l; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python tools/bench_analysis.py > $O/analysis.jsonl 2> $O/analysis.err
timeout 600 ncu --set full --clock-control none -k regex:"rdf_hist_kernel|time_corr" -c 3 -o $O/ncu_analysis python tools/bench_analysis.py --cpu-seconds 0.5 --repeat 1 > $O/ncu_analysis.log 2>&1
ls $O
