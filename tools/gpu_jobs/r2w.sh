set -x
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g e2e_pinned=%.4g kernel_ms=%.3f F=%d wave=%d" % (d["value"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["roofline"]["kernel_ms"], d["run"]["frames_per_step_per_gpu"], d["run"]["wave_frames"]))'
timeout 150 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "C4"
POLAR_B200_FORCE_SPLIT=0 timeout 150 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "C4 single"
POLAR_B200_WARPS_PER_CTA=4 timeout 150 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "C4 W=4"
POLAR_B200_WARPS_PER_CTA=1 timeout 150 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "C4 W=1"
POLAR_B200_CTAS_PER_SM=4 timeout 150 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "C4 4 CTAs/SM"
timeout 200 python tools/bench_kinds.py --only Fast 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('%-45s %-13s %.4g' % (d.get('shape', d.get('name','?')), d['kernel'], d['frames_per_s']))
"
