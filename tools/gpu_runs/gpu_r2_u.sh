#!/bin/bash
# ncu --set full of the list-build kernel at N = 2^20 (first force evaluation = one full rebuild).
O=gpurun_out/r2u; mkdir -p $O
LJMD_NL_NOGRAPH=1 STEPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:nl_build_kernel -c 1 -o $O/ncu_nl_build python tools/quick_nlist.py cfg4_nlist > $O/ncu_nl_build.log 2>&1
echo "rc=$?" >> $O/ncu_nl_build.log
tail -3 $O/ncu_nl_build.log; ls -la $O
