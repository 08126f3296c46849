#!/bin/bash
# round 2, job c: GPU suite with the new host paths / toys / reference callers / checkpoint; new bench.py at the driver's settings
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02c_gputests.log 2>&1
tail -n 30 gpurun_out/r02c_gputests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_driver.json 2> gpurun_out/r02c_bench_driver.err
timeout 600 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --no-toys > gpurun_out/r02c_bench_1000.json 2> gpurun_out/r02c_bench_1000.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02c_bench_reference.json 2> gpurun_out/r02c_bench_reference.err
cat gpurun_out/r02c_bench_driver.json gpurun_out/r02c_bench_1000.json gpurun_out/r02c_bench_reference.json
tail -n 5 gpurun_out/r02c_bench_driver.err gpurun_out/r02c_bench_reference.err
