#!/bin/bash
# round 2, 8-GPU call: real 8-process slab test + strong-scaling variants at N = 2^20
N=8
O=gpurun_out/r2g_n$N; mkdir -p $O
P=$PWD/molecular-dynamics-simulation---lennard-jones-monoatomic-fluid_b200
nvidia-smi -L > $O/gpus.txt
LJMD_SLAB_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_multi_process.py -x -q -k "8]" > $O/pytest_mp.log 2>&1; echo "rc=$?" >> $O/pytest_mp.log
run() { # name, gpus, env...
  name=$1; n=$2; shift; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $n --steps 200 --warmup 20 > $O/bench_${name}_n$n.json 2> $O/bench_${name}_n$n.err
  echo "$name n=$n rc=$?" >> $O/runs.log
}
run base 8 LJMD_X=0
run pub 8 LJMD_LIB_PATH=$P/lib/libljmd_pub.so
run pipe1 8 LJMD_NL_PIPE=1
run pub_skin04 8 LJMD_LIB_PATH=$P/lib/libljmd_pub.so LJMD_NL_SKIN=0.4
run pub 4 LJMD_LIB_PATH=$P/lib/libljmd_pub.so
ls -la $O
