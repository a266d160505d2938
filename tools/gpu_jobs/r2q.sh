set -x
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g e2e_pinned=%.4g kernel_ms=%.3f F=%d wave=%d" % (d["value"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["roofline"]["kernel_ms"], d["run"]["frames_per_step_per_gpu"], d["run"]["wave_frames"]))'
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "default"
POLAR_B200_FORCE_SPLIT=0 timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "single"
for c in C2 C3 C4; do timeout 150 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "$c"; done
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sweep.py -m gpu -x -q --timeout 120 2>&1 | tail -4
