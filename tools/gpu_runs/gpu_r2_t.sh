#!/bin/bash
# Drop-in call at N = 2^20: chunk-pipelined upload / drift / position download (LJMD_COMPAT_OVERLAP=2) vs =1 vs =0.
O=gpurun_out/r2t; mkdir -p $O
timeout 240 python -m pytest tests/test_zz_gpu_ab_switches.py tests/test_gpu_nlist.py -k "overlap or page_locked or drop_in" -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
for m in 2 1; do
LJMD_COMPAT_OVERLAP=$m timeout 120 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu --long-steps 0 > $O/bench_overlap$m.json 2> $O/bench_overlap$m.err; echo "rc=$?" >> $O/bench_overlap$m.err
done
LJMD_COMPAT_CHUNKS=8 timeout 120 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu --long-steps 0 > $O/bench_overlap2_c8.json 2> $O/bench_overlap2_c8.err
tail -3 $O/pytest.log; python - <<'P'
import json
for f in ("bench_overlap2", "bench_overlap1", "bench_overlap2_c8"):
    try:
        d = json.loads(open("gpurun_out/r2t/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["e2e"]["ms_per_step"], d["e2e"]["value"], d["ms_per_step"])
    except Exception as e:
        print(f, "failed", e)
P
