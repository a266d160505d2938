#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_random_configs.py -x -q -m gpu 2>&1 | tail -3
for s in 11 12; do timeout 300 python tests/fuzz_parity.py $s 2>&1 | tail -1; done
python tools/bench_kinds.py --only "C4" 2>&1 | tail -1 | cut -c1-220
python tools/bench_kinds.py --only "NS " 2>&1 | tail -1 | cut -c1-220
python tools/bench_kinds.py --only "float FastSCL" 2>&1 | tail -1 | cut -c1-220
