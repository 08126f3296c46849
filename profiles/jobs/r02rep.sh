#!/bin/bash
# round 2: the GPU suite three times in a row on one box (flakiness check)
for i in 1 2 3; do
  timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02rep_$i.log 2>&1; echo "run $i rc=$?"; tail -n 1 gpurun_out/r02rep_$i.log
done
