#!/bin/bash
# submit.sh <timeout_s> <job script> [gpus]: gpurun with retries while the pod answers "busy" (exit code 3)
T=$1; JOB=$2; G=${3:-1}
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "bash $JOB"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $JOB"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
