set -x
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:path_warp -s 1 -c 1 -o gpurun_out/prof_r2_fscl python tools/bench_kinds.py --only "float SCL" --reps 1 > gpurun_out/r2ze_fscl.log 2>&1; tail -2 gpurun_out/r2ze_fscl.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:path_warp -s 1 -c 1 -o gpurun_out/prof_r2_c5 python tools/bench_kinds.py --only "C5" --reps 1 > gpurun_out/r2ze_c5.log 2>&1; tail -2 gpurun_out/r2ze_c5.log
ls -la gpurun_out/prof_r2_*.ncu-rep
