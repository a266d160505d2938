#!/bin/bash
# Round-end check on a B200 box: all GPU tests, smoke(), both bench arms and the per-shape survey (writes under gpurun_out/).
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err; tail -2 gpurun_out/bench_r1_g.err; cat gpurun_out/bench_r1_g.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_g.json 2> gpurun_out/bench_ref_g.err; tail -2 gpurun_out/bench_ref_g.err; cat gpurun_out/bench_ref_g.json
python tools/bench_kinds.py > gpurun_out/kinds_r1_g.jsonl 2>/dev/null; wc -l gpurun_out/kinds_r1_g.jsonl
