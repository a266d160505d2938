set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 90 -k "staging or set_devices or full_residency or survives" > gpurun_out/r2l_tests.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2l_tests.log
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g e2e_pinned=%.4g kernel_ms=%.3f F=%d" % (d["value"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["roofline"]["kernel_ms"], d["run"]["frames_per_step_per_gpu"]))'
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --verbose > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"; tail -6 gpurun_out/r2l_bench.err; python -c "$P" default < gpurun_out/r2l_bench.json
for f in 1 2 3; do POLAR_B200_KFLAGS=$f timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "KFLAGS=$f"; done
POLAR_B200_FORCE_SPLIT=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "single-launch"
timeout 300 python tools/gpu_jobs/probe2.py knobs > gpurun_out/r2l_knobs.log 2>&1; cat gpurun_out/r2l_knobs.log
