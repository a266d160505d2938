set -x
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g e2e_pinned=%.4g kernel_ms=%.3f F=%d wave=%d" % (d["value"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["roofline"]["kernel_ms"], d["run"]["frames_per_step_per_gpu"], d["run"]["wave_frames"]))'
for cfg in "4 2" "4 3" "4 4" "3 5" "2 6" "2 4"; do set -- $cfg; POLAR_B200_WARPS_PER_CTA=$1 POLAR_B200_CTAS_PER_SM=$2 timeout 150 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "C4 W=$1 CTAs=$2"; done
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 132608 2>/dev/null | python -c "$P" "NS pieces"
for st in 0 80 160; do POLAR_B200_STAGGER_US=$st POLAR_B200_FORCE_SPLIT=0 timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 132608 2>/dev/null | python -c "$P" "NS single stagger_us=$st"; done
POLAR_B200_STAGGER_US=160 POLAR_B200_FORCE_SPLIT=0 timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "NS single stagger_us=160 32 waves"
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "NS pieces 32 waves"
