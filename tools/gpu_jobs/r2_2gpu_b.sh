set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 150 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_2gpu_b_NS.json 2> gpurun_out/r2_2gpu_b_NS.err; echo "NS rc=$?"; cat gpurun_out/r2_2gpu_b_NS.json | cut -c1-1500; tail -3 gpurun_out/r2_2gpu_b_NS.err
