set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "set_devices" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_2gpu_NS.json 2> gpurun_out/r2_2gpu_NS.err; echo "NS rc=$?"; cut -c1-250 gpurun_out/r2_2gpu_NS.json
timeout 200 $TR bench.py --gpus 2 --config C4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_2gpu_C4.json 2> gpurun_out/r2_2gpu_C4.err; echo "C4 rc=$?"; cut -c1-250 gpurun_out/r2_2gpu_C4.json
timeout 200 $TR bench.py --gpus 2 --config C5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_2gpu_C5.json 2> gpurun_out/r2_2gpu_C5.err; echo "C5 rc=$?"; cut -c1-250 gpurun_out/r2_2gpu_C5.json
timeout 120 python - <<'PY'
# one decode() call sharded over both GPUs by the library (pd_set_devices), no torchrun
import time, numpy as np, sys
sys.path.insert(0, "tests")
import bench, quantized_decoder_polar_codes_b200 as q
cfg = bench.CONFIGS["NS"]
kw, x, _ = bench.make_workload(cfg, 8192, 1)
dec = q.SCLLUTDecoder(**kw)
X = np.tile(x, (64, 1)).astype(np.uint8)
ref = dec.decode(X[:8192])
for ids in ([0], [0, 1]):
    dec.set_devices(ids)
    dec.decode(X)
    t0 = time.perf_counter(); out = dec.decode(X); dt = time.perf_counter() - t0
    print("set_devices", ids, "%.4g frames/s" % (X.shape[0] / dt), "equal", bool((out[:8192] == ref).all() and (out[-8192:] == ref).all()))
PY
