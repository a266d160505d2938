set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,temperature.gpu --format=csv,noheader
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g e2e_pinned=%.4g kernel_ms=%.3f F=%d clocks=%s" % (d["value"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["roofline"]["kernel_ms"], d["run"]["frames_per_step_per_gpu"], d["clocks"]))'
export POLAR_B200_RING=private
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 90 -k "staging or set_devices or full_residency or survives" > gpurun_out/r2m_tests.log 2>&1; echo "pytest(private) rc=$?"; tail -6 gpurun_out/r2m_tests.log
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --verbose > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench(private) rc=$?"; tail -4 gpurun_out/r2m_bench.err; python -c "$P" private < gpurun_out/r2m_bench.json
POLAR_B200_FORCE_SPLIT=0 timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "private single-launch"
POLAR_B200_RING=shared POLAR_B200_NO_TMA=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_notma.json 2> gpurun_out/r2m_bench_notma.err; echo "bench(shared,NO_TMA) rc=$?"; python -c "$P" shared-notma < gpurun_out/r2m_bench_notma.json
timeout 500 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r2m_full.log 2>&1; echo "full pytest(private) rc=$?"; tail -6 gpurun_out/r2m_full.log
