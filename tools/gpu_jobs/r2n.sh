set -x
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g e2e_pinned=%.4g kernel_ms=%.3f F=%d" % (d["value"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["roofline"]["kernel_ms"], d["run"]["frames_per_step_per_gpu"]))'
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench(private) rc=$?"; python -c "$P" private < gpurun_out/r2n_bench.json
POLAR_B200_FORCE_SPLIT=0 timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "private single-launch"
POLAR_B200_RING=shared POLAR_B200_NO_TMA=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" shared-notma
POLAR_B200_RING=shared POLAR_B200_NO_TMA=3 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" shared-tma-elect-cta
POLAR_B200_RING=shared POLAR_B200_NO_TMA=3 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 60 -k "staging or set_devices or full_residency or survives" 2>&1 | tail -3
timeout 500 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r2n_full.log 2>&1; echo "full pytest(private) rc=$?"; tail -6 gpurun_out/r2n_full.log
