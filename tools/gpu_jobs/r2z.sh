set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2z_smoke.log
timeout 700 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/r2z_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2z_tests.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2z_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2z_launches.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:scl_lut -s 4 -c 6 --csv --log-file gpurun_out/r2z_traffic.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_traffic.log 2>&1; echo "traffic rc=$?"; tail -8 gpurun_out/r2z_traffic.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scl_lut -s 4 -c 1 -o gpurun_out/prof_r2z python bench.py --steps 3 --warmup 1 --no-cpu-baseline --batch 132608 > gpurun_out/r2z_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/bench_kinds.py > gpurun_out/r2z_kinds.jsonl 2> gpurun_out/r2z_kinds.err; echo "kinds rc=$?"
ls -la gpurun_out | tail -12
