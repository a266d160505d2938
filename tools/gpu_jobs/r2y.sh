set -x
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value=%.4g e2e=%.4g e2e_pinned=%.4g kernel_ms=%.3f F=%d wave=%d" % (d["value"], d["e2e"]["value"], d["e2e_pinned"]["value"], d["roofline"]["kernel_ms"], d["run"]["frames_per_step_per_gpu"], d["run"]["wave_frames"]))'
cp quantized_decoder_polar_codes_b200/libpolar_b200.so /tmp/lib_orig.so
for v in v0 v2 v3 v5; do
cp gpu_variants/libpolar_b200_$v.so quantized_decoder_polar_codes_b200/libpolar_b200.so
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "NS $v pieces"
POLAR_B200_FORCE_SPLIT=0 timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "NS $v single"
done
cp /tmp/lib_orig.so quantized_decoder_polar_codes_b200/libpolar_b200.so
timeout 150 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "C4 (12 warps default)"
POLAR_B200_FORCE_SPLIT=0 timeout 150 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P" "C4 single"
