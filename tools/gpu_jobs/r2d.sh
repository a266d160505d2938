set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1 || { echo SMOKE_FAILED; tail -20 gpurun_out/r2d_smoke.log; exit 1; }
tail -5 gpurun_out/r2d_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log; tail -15 gpurun_out/r2d_tests.log
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g kernel_ms=%.3f F=%d wave=%s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["config"]["frames_per_step_per_gpu"], d["config"].get("wave_frames")))'
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; cat gpurun_out/r2d_bench.json | python -c "$P" "default"
POLAR_B200_FORCE_SPLIT=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "single"
POLAR_B200_WARPS_PER_CTA=3 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "W=3"
POLAR_B200_WARPS_PER_CTA=2 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "W=2"
for s in "SC-LUT N=1024" "CASCL-LUT" "C4 " "FastSC-LUT" "float SCL" "float FastSCL" "C5 " "Lloyd"; do
  date +%T; timeout 150 python tools/bench_kinds.py --only "$s" >> gpurun_out/r2d_kinds.jsonl 2>> gpurun_out/r2d_kinds.err || echo "KINDS FAILED/TIMEOUT: $s"
done
python -c "
import json
for l in open('gpurun_out/r2d_kinds.jsonl'):
    d=json.loads(l); print(d.get('shape'), d.get('kernel'), '%.3g'%d.get('frames_per_s',0))
"
date +%T
timeout 900 python tools/make_mmi_n1024.py gpurun_out/mmi_n1024_q16_3dB.npz 2>&1 | tail -3
date +%T
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scl_lut -s 8 -c 1 -o gpurun_out/prof_r2d python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_ncu.log 2>&1
ls -la gpurun_out/
