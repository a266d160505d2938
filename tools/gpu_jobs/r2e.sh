set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g kernel_ms=%.3f F=%d wave=%s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["config"]["frames_per_step_per_gpu"], d["config"].get("wave_frames")))'
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "default(hint)"
timeout 900 python tools/gpu_jobs/probe_l1.py 2>&1 | tail -40
