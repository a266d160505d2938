set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 700 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/r2zc_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2zc_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2zc_bench.json 2> gpurun_out/r2zc_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/r2zc_bench.json
for c in C2 C3; do timeout 200 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2zc_$c.json 2>gpurun_out/r2zc_$c.err; echo "$c rc=$?"; cut -c1-200 gpurun_out/r2zc_$c.json; done
