set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1 || { echo SMOKE_FAILED; tail -20 gpurun_out/r2b_smoke.log; exit 1; }
tail -6 gpurun_out/r2b_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log; tail -4 gpurun_out/r2b_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['e2e']['value'], d['roofline']['kernel_ms'])"
for w in 3 4 5; do POLAR_B200_WARPS_PER_CTA=$w timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('W=$w', d['value'], d['roofline']['kernel_ms'])"; done
timeout 300 ncu --metrics smsp__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:scl_lut -s 8 -c 4 --csv --log-file gpurun_out/r2b_ncu_inst.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_ncu.log 2>&1
tail -30 gpurun_out/r2b_ncu_inst.csv | cut -c1-300
