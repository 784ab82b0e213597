set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1: geo %.4f ms  %.0f GB/s  frac %.3f | step %.4f ms value %.2fM e2e %.2fM' % (r['avg_launch_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))
"; }
timeout 300 python bench.py --no-cpu-baseline --steps 40 2> gpurun_out/h.err | summ streamed
timeout 300 python bench.py --no-cpu-baseline --steps 40 2>> gpurun_out/h.err | summ streamed_again
export TSPN_NVCC_EXTRA="-DTSPN_GEO_STREAMED=0"
python -m tspn_b200.build --force > /dev/null 2> gpurun_out/sweep_build.err || tail -5 gpurun_out/sweep_build.err
timeout 300 python bench.py --no-cpu-baseline --steps 40 2>> gpurun_out/h.err | summ burst
unset TSPN_NVCC_EXTRA
tail -5 gpurun_out/h.err
