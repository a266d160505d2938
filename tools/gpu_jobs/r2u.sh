set -x
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g e2e_pinned=%.4g kernel_ms=%.3f F=%d wave=%d" % (d["value"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["roofline"]["kernel_ms"], d["run"]["frames_per_step_per_gpu"], d["run"]["wave_frames"]))'
for w in 1 2 4; do
POLAR_B200_PIECE_WAVES=$w timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "NS piece_waves=$w"
done
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 131072 2>/dev/null | python -c "$P" "NS B=131072"
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 1000000 2>/dev/null | python -c "$P" "NS B=1e6"
for c in C1 C3 C4 C5; do for w in 1 2; do POLAR_B200_PIECE_WAVES=$w timeout 150 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "$c piece_waves=$w"; done; done
