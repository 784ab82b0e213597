# A/B of the pair-geometry kernel forms + ring-depth sweep of the persistent one (GPU box).
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/ab_pytest.log
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1: geo %.4f ms  %.0f GB/s  frac %.3f | step %.4f ms value %.2fM e2e %.2fM' % (r['avg_launch_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))
"; }
timeout 300 python bench.py --no-cpu-baseline --steps 30 2> gpurun_out/ab.err | tee gpurun_out/ab_per_item.json | summ per_item
TSPN_GEO_PERSISTENT=1 timeout 300 python bench.py --no-cpu-baseline --steps 30 2>> gpurun_out/ab.err | tee gpurun_out/ab_persistent.json | summ persistent
for ring in 3 2; do
  export TSPN_NVCC_EXTRA="-DTSPN_GEO_PRING=$ring"
  python -m tspn_b200.build --force > /dev/null 2> gpurun_out/sweep_build.err || { echo "ring $ring: build failed"; tail -5 gpurun_out/sweep_build.err; continue; }
  timeout 300 python bench.py --no-cpu-baseline --steps 30 2>> gpurun_out/ab.err | summ persistent_ring$ring
done
unset TSPN_NVCC_EXTRA
tail -5 gpurun_out/ab.err
