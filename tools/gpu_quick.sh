set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/q_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline 2> gpurun_out/q_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.2fM e2e %.2fM step %.4f geo %.4f frac %.3f alone %.4f / %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['alone']['avg_launch_ms'], r['alone']['frac']))
"
done
