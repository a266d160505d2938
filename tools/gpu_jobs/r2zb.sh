set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sim.py -m gpu -x -q --timeout 180 -k "error_counters or sim or north_star" 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2zb_bench.json 2> gpurun_out/r2zb_bench.err; echo "bench rc=$?"; cat gpurun_out/r2zb_bench.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2zb_ref.json 2> gpurun_out/r2zb_ref.err; echo "ref rc=$?"; cat gpurun_out/r2zb_ref.json
for c in C1 C2 C3 C4 C5; do timeout 200 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2zb_$c.json 2>gpurun_out/r2zb_$c.err; echo "$c rc=$?"; cut -c1-260 gpurun_out/r2zb_$c.json; done
