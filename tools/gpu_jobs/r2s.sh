set -x
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g e2e_pinned=%.4g kernel_ms=%.3f F=%d wave=%d" % (d["value"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["roofline"]["kernel_ms"], d["run"]["frames_per_step_per_gpu"], d["run"]["wave_frames"]))'
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "NS"
for c in C4 C2 C3; do timeout 150 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "$c"; done
POLAR_B200_SMEM_LEVEL_WORDS=8 timeout 150 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "C4 level_words=8"
POLAR_B200_SMEM_LEVEL_WORDS=2 timeout 150 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "C4 level_words=2"
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "staging" 2>&1 | tail -4
