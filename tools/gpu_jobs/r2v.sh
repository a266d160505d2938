set -x
mkdir -p gpurun_out
timeout 400 python tools/bench_kinds.py > gpurun_out/r2_kinds.jsonl 2> gpurun_out/r2_kinds.err; echo "kinds rc=$?"; python -c "
import json
for l in open('gpurun_out/r2_kinds.jsonl'):
    d=json.loads(l); print('%-45s %-13s %.4g' % (d.get('shape', d.get('name','?')), d['kernel'], d['frames_per_s']))
"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench.json 2>gpurun_out/r2v_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2v_bench.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scl_lut -s 6 -c 1 -o gpurun_out/prof_r2v_c4 python bench.py --config C4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2v_ncu.log 2>&1
tail -2 gpurun_out/r2v_ncu.log
