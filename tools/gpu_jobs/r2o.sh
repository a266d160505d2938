set -x
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g e2e_pinned=%.4g kernel_ms=%.3f F=%d" % (d["value"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["roofline"]["kernel_ms"], d["run"]["frames_per_step_per_gpu"]))'
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench(private) rc=$?"; python -c "$P" private < gpurun_out/r2o_bench.json
POLAR_B200_RING=shared POLAR_B200_NO_TMA=3 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" shared-tma-elect-cta
POLAR_B200_RING=shared POLAR_B200_NO_TMA=3 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 60 -k "staging or set_devices or full_residency or survives" 2>&1 | tail -3
POLAR_B200_FORCE_SPLIT=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:scl_lut -s 4 -c 1 -o gpurun_out/prof_r2o python bench.py --steps 2 --warmup 1 --no-cpu-baseline --batch 28416 > gpurun_out/r2o_ncu.log 2>&1
tail -3 gpurun_out/r2o_ncu.log
ls -la gpurun_out | tail -5
