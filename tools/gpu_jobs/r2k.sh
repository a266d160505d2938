set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 90 -k "staging or set_devices or full_residency or survives" 2>&1 | tail -15
POLAR_B200_TRACE=1 POLAR_B200_PINNED_RESULT_MIN=1000000000000 timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --verbose > gpurun_out/r2k_bench_nopin.json 2> gpurun_out/r2k_bench_nopin.err; echo "bench(nopin) rc=$?"; tail -40 gpurun_out/r2k_bench_nopin.err; cut -c1-200 gpurun_out/r2k_bench_nopin.json
POLAR_B200_TRACE=1 timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --verbose > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; tail -40 gpurun_out/r2k_bench.err; cut -c1-200 gpurun_out/r2k_bench.json
