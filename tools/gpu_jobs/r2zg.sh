set -x
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 80 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 60 -k "north_star or full_residency" 2>&1 | tail -2
timeout 60 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:scl_lut -s 3 -c 3 --csv --log-file gpurun_out/r2zg_traffic.csv python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r2zg_traffic.log 2>&1; grep -c scl_lut gpurun_out/r2zg_traffic.csv; grep "dram__bytes" gpurun_out/r2zg_traffic.csv | cut -d, -f12- | head -6
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g kernel_ms=%.3f equal=%s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["run"]["e2e_outputs_equal_device_output"]))'
timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "discard"
POLAR_B200_KDEBUG=2 timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "no-discard"
