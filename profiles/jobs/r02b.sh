#!/bin/bash
# round 2, job b: full GPU test-suite on the fixed-point-carrier kernel, refill-vote ablation, determinism, big parity report, bench
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02b_gputests.log 2>&1
export NV=$PWD/sac-agent_b200/libboatenv_novote.so
BOATENV_LIBRARY=$NV timeout 600 python -m pytest tests/test_gpu_benchmark_regime.py -q > gpurun_out/r02b_novote.log 2>&1
BOATENV_LIBRARY=$NV timeout 300 python profiles/determinism_check.py 20 > gpurun_out/r02b_novote_det.log 2>&1
timeout 300 python profiles/determinism_check.py 20 > gpurun_out/r02b_vote_det.log 2>&1
timeout 900 python profiles/parity_report.py big 32768 2000 1 2 3 4 5 > gpurun_out/r02b_parity_big.jsonl 2> gpurun_out/r02b_parity_big.err
timeout 600 python bench.py --steps 300 --warmup 600 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
tail -n 5 gpurun_out/r02b_gputests.log gpurun_out/r02b_novote.log gpurun_out/r02b_novote_det.log gpurun_out/r02b_vote_det.log
cat gpurun_out/r02b_parity_big.jsonl gpurun_out/r02b_bench.json
