set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2zf_bench.json 2> gpurun_out/r2zf_bench.err; echo "bench rc=$?"; cat gpurun_out/r2zf_bench.json; tail -3 gpurun_out/r2zf_bench.err
for c in C1 C4; do timeout 200 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2zf_$c.json 2>gpurun_out/r2zf_$c.err; echo "$c rc=$?"; python -c "
import json,sys
d=json.load(open('gpurun_out/r2zf_$c.json')); print('$c', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'api %.4g'%d['e2e_api']['value'], 'cabi %.4g'%d['e2e_cabi']['value'])
"; tail -2 gpurun_out/r2zf_$c.err; done
