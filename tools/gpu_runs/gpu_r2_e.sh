#!/bin/bash
# round 2, multi-GPU call (run with gpurun --gpus N): real multi-process slab tests + bench variants
N=${1:-2}
O=gpurun_out/r2e_n$N; mkdir -p $O
P=$PWD/molecular-dynamics-simulation---lennard-jones-monoatomic-fluid_b200
nvidia-smi -L > $O/gpus.txt
LJMD_SLAB_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_multi_process.py -x -q -k "$2" > $O/pytest_mp.log 2>&1; echo "rc=$?" >> $O/pytest_mp.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --steps 200 --warmup 20 > $O/bench_$name.json 2> $O/bench_$name.err
  echo "$name rc=$?" >> $O/runs.log
}
run base LJMD_X=0
run pipe0 LJMD_NL_PIPE=0
run sh1024_pipe0 LJMD_NL_PIPE=0 LJMD_LIB_PATH=$P/lib/libljmd_sh1024.so
run pub_pipe0 LJMD_NL_PIPE=0 LJMD_LIB_PATH=$P/lib/libljmd_pub.so
run pub_pipe1 LJMD_NL_PIPE=1 LJMD_LIB_PATH=$P/lib/libljmd_pub.so
ls -la $O
