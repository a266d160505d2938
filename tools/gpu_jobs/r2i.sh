set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python tools/gpu_jobs/probe2.py flags 2>&1 | tail -10
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g e2e_pinned=%.4g kernel_ms=%.3f F=%d" % (d["value"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["roofline"]["kernel_ms"], d["run"]["frames_per_step_per_gpu"]))'
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2i_bench.err | tee gpurun_out/r2i_bench.json | python -c "$P" "NS"
POLAR_B200_NO_TMA=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "NS-no-tma"
POLAR_B200_NO_TMA=4 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "NS-tma-alive"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scl_lut -s 8 -c 1 -o gpurun_out/prof_r2i python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2i_ncu.log 2>&1
ls -la gpurun_out | tail -5
