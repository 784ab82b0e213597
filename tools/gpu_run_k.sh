# side-branch priority A/B + timeline of the step (CUPTI)
set -x
mkdir -p gpurun_out
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1: geo %.4f ms  %.0f GB/s  frac %.3f | step %.4f ms value %.2fM e2e %.2fM' % (r['avg_launch_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))
"; }
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/k_pytest.log
TSPN_SIDE_PRIORITY=0 timeout 300 python bench.py --no-cpu-baseline --steps 30 2> gpurun_out/k_p0.err | tee gpurun_out/k_p0.json | summ prio0
TSPN_SIDE_PRIORITY=-1 timeout 300 python bench.py --no-cpu-baseline --steps 30 2> gpurun_out/k_p1.err | tee gpurun_out/k_p1.json | summ prio-1
TSPN_SIDE_PRIORITY=0 timeout 300 python tools/trace_step.py --steps 2 > gpurun_out/k_trace_p0.txt 2> gpurun_out/k_trace_p0.err; tail -3 gpurun_out/k_trace_p0.err
TSPN_SIDE_PRIORITY=-1 timeout 300 python tools/trace_step.py --steps 2 > gpurun_out/k_trace_p1.txt 2> gpurun_out/k_trace_p1.err; tail -3 gpurun_out/k_trace_p1.err
cat gpurun_out/k_trace_p1.txt | tail -40
