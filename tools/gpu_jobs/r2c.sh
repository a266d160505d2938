set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1 || { echo SMOKE_FAILED; tail -20 gpurun_out/r2c_smoke.log; exit 1; }
tail -5 gpurun_out/r2c_smoke.log
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g kernel_ms=%.3f F=%d wave=%s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["config"]["frames_per_step_per_gpu"], d["config"].get("wave_frames")))'
for w in 4 3 6 5; do
  POLAR_B200_WARPS_PER_CTA=$w timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "W=$w single"
done
POLAR_B200_WARPS_PER_CTA=4 POLAR_B200_FORCE_SPLIT=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "W=4 split"
POLAR_B200_WARPS_PER_CTA=4 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 131072 2>/dev/null | python -c "$P" "W=4 batch131072(split)"
POLAR_B200_WARPS_PER_CTA=4 timeout 600 python tools/bench_kinds.py > gpurun_out/r2c_kinds.jsonl 2> gpurun_out/r2c_kinds.err
python -c "
import json
for l in open('gpurun_out/r2c_kinds.jsonl'):
    d=json.loads(l); print(d.get('shape'), d.get('kernel'), '%.3g'%d.get('frames_per_s',0))
"
