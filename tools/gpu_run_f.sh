set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/f_pytest.log
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1: geo %.4f ms  %.0f GB/s  frac %.3f | step %.4f ms value %.2fM e2e %.2fM' % (r['avg_launch_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))
"; }
timeout 300 python bench.py --no-cpu-baseline --steps 40 2> gpurun_out/f.err | tee gpurun_out/f_bench_dense.json | summ dense
TSPN_GEO_SPARSE=1 timeout 300 python bench.py --no-cpu-baseline --steps 40 2>> gpurun_out/f.err | tee gpurun_out/f_bench_sparse.json | summ sparse
timeout 300 python bench.py --no-cpu-baseline --steps 40 2>> gpurun_out/f.err | summ dense_again
tail -5 gpurun_out/f.err
