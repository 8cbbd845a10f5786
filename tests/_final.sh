#!/bin/bash
# full GPU suite, default bench line, then the ncu captures of the same build
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -c 600 gpurun_out/r02_bench_1gpu.json
bash tests/_ncu.sh > gpurun_out/r02_ncu_sh.log 2>&1; tail -3 gpurun_out/r02_ncu_sh.log
