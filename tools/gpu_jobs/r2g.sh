set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python tools/gpu_jobs/probe2.py verify 2>&1 | tail -12
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g e2e_pinned=%.4g kernel_ms=%.3f F=%d" % (d["value"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["roofline"]["kernel_ms"], d["run"]["frames_per_step_per_gpu"]))'
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2g_bench.err | tee gpurun_out/r2g_bench.json | python -c "$P" "NS"
tail -3 gpurun_out/r2g_bench.err
