set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python tools/gpu_jobs/probe2.py knobs 2>&1 | tail -12
timeout 600 python tools/gpu_jobs/probe2.py lists 2>&1 | tail -12
timeout 400 python tools/gpu_jobs/probe2.py memcheck > gpurun_out/r2f_memcheck.log 2>&1; tail -40 gpurun_out/r2f_memcheck.log
nvidia-smi --query-gpu=name,memory.used --format=csv
