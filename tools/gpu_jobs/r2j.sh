set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --verbose > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; cat gpurun_out/r2j_bench.err | tail -8; cut -c1-300 gpurun_out/r2j_bench.json
timeout 400 python tools/gpu_jobs/probe2.py flags > gpurun_out/r2j_flags.log 2>&1; cat gpurun_out/r2j_flags.log
