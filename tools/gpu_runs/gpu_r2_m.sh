#!/bin/bash
# round 2, third 8-GPU call: final defaults (publish in kernel, slab-local rebuild, prepared graphs) -- bitwise tests at 2/4/8 ranks,
# driver-style scaling runs (--steps 20 --warmup 5), a 200-step run and the per-step timeline at 8
O=gpurun_out/r2n_n8; mkdir -p $O
LJMD_SLAB_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_multi_process.py -x -q > $O/pytest_mp.log 2>&1; echo "rc=$?" >> $O/pytest_mp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/slab_timeline.py 96 > $O/timeline_n8.txt 2>&1
run() { name=$1; n=$2; st=$3; wu=$4; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $n --steps $st --warmup $wu > $O/bench_${name}_n$n.json 2> $O/bench_${name}_n$n.err; echo "$name n=$n rc=$?" >> $O/runs.log; }
run driver 8 20 5
run driver 4 20 5
run driver 2 20 5
run long 8 200 20
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-secondary > $O/bench_driver_n1.json 2> $O/bench_driver_n1.err
ls $O
