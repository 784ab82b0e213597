# ring depth of the pair kernel x survivor path (co-residency needs shared memory left over)
set -x
mkdir -p gpurun_out
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1: geo %.4f ms  %.0f GB/s  frac %.3f | step %.4f ms value %.2fM e2e %.2fM' % (r['avg_launch_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))
"; }
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 30 2> gpurun_out/n_$tag.err | tee gpurun_out/n_$tag.json | summ $tag || tail -3 gpurun_out/n_$tag.err; }
for ring in 2 3; do
  export TSPN_NVCC_EXTRA="-DTSPN_GEO_RING=$ring"
  python -m tspn_b200.build --force > /dev/null 2> gpurun_out/n_build_$ring.err || { echo "ring $ring build failed"; tail -5 gpurun_out/n_build_$ring.err; continue; }
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor.py -x -q -k "pair_geometry or kernel_shapes or survivor" > gpurun_out/n_parity_$ring.log 2>&1; echo "ring $ring parity rc=$?"; tail -2 gpurun_out/n_parity_$ring.log
  run ring${ring}_surv0 TSPN_SURVIVOR_PATH=0
  run ring${ring}_surv1 TSPN_SURVIVOR_PATH=1
  run ring${ring}_surv1_p0 TSPN_SURVIVOR_PATH=1 TSPN_SIDE_PRIORITY=0
  TSPN_SURVIVOR_PATH=1 timeout 300 python tools/trace_step.py --steps 2 > gpurun_out/n_trace_$ring.txt 2> gpurun_out/n_trace_$ring.err
  tail -18 gpurun_out/n_trace_$ring.txt
done
