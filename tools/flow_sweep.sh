#!/bin/bash
# usage: tools/flow_sweep.sh "VAR=val VAR2=val" "..." ; each argument is one environment for tools/profile_step.py
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 300 python tools/profile_step.py --settle 1500 --steps 5 2>&1 | grep -E "ms/step|flow solver" | sed -e "s/'integrate_v'.*'solve_kernel'/solve_kernel/"
done
