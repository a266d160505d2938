set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python tools/gpu_jobs/probe2.py knobs 2>&1 | tail -8
timeout 600 python tools/gpu_jobs/probe2.py lists 2>&1 | tail -8
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g e2e_pinned=%.4g kernel_ms=%.3f F=%d" % (d["value"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["roofline"]["kernel_ms"], d["run"]["frames_per_step_per_gpu"]))'
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2h_bench.err | tee gpurun_out/r2h_bench.json | python -c "$P" "NS"
for s in "SC-LUT N=1024" "FastSC-LUT" "C4 " "C2 "; do timeout 150 python tools/bench_kinds.py --only "$s" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d.get('shape'), d.get('kernel'), '%.3g'%d.get('frames_per_s',0))" || echo "KINDS FAILED: $s"; done
