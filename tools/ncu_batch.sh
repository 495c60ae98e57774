#!/bin/bash
# one ncu --set full capture of the batched-worlds solve kernel; usage: tools/ncu_batch.sh <kernel regex> <out name>
K=${1:-k_solve_groups}; O=${2:-prof_batch}
timeout 200 python tools/profile_batch.py 4096 300 1 > gpurun_out/plain_$O.log 2>&1 &&
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/$O python tools/profile_batch.py 4096 300 1 > gpurun_out/ncu_$O.log 2>&1
tail -2 gpurun_out/ncu_$O.log
