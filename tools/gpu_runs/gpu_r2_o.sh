#!/bin/bash
# final 1-GPU check of the round: the driver's own sequence (tests, smoke, reference arm, default bench) + one ncu capture of the step kernel
O=gpurun_out/r2o; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_default.csv python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-secondary > $O/launches_default.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nl_step_kernel -c 1 -s 3 -o $O/ncu_nl_step python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu --no-secondary > $O/ncu_nl_step.log 2>&1
ls $O
