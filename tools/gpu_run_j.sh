# One GPU call: full parity suite + bench on the default build, then A/B of the pair-geometry kernel variants
# (object-group size, 1024 threads x 2 frames, bulk stores from shared memory), each with a parity subset.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/j_pytest.log
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1: geo %.4f ms  %.0f GB/s  frac %.3f | step %.4f ms value %.2fM e2e %.2fM' % (r['avg_launch_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))
"; }
timeout 300 python bench.py --no-cpu-baseline --steps 30 2> gpurun_out/j_default.err | tee gpurun_out/j_default.json | summ default
tail -3 gpurun_out/j_default.err
i=0
for flags in "-DTSPN_GEO_OBJ_GROUP=32" "-DTSPN_GEO_OBJ_GROUP=16" "-DTSPN_GEO_WIDE=1" "-DTSPN_GEO_TMA_STORE=1" "-DTSPN_GEO_TMA_STORE=2" "-DTSPN_GEO_WIDE=1 -DTSPN_GEO_TMA_STORE=1"; do
  i=$((i+1))
  export TSPN_NVCC_EXTRA="$flags"
  python -m tspn_b200.build --force > /dev/null 2> gpurun_out/j_build_$i.err || { echo "[$flags] build failed"; tail -5 gpurun_out/j_build_$i.err; continue; }
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "pair_geometry or kernel_shapes" > gpurun_out/j_parity_$i.log 2>&1; echo "[$flags] parity rc=$?"
  tail -3 gpurun_out/j_parity_$i.log
  timeout 300 python bench.py --no-cpu-baseline --steps 30 2> gpurun_out/j_var_$i.err | tee gpurun_out/j_var_$i.json | summ "[$flags]" || tail -3 gpurun_out/j_var_$i.err
done
unset TSPN_NVCC_EXTRA
ls gpurun_out | head -50
