# survivor path: parity + A/B (survivor path on/off, side priority) + timeline
set -x
mkdir -p gpurun_out
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1: geo %.4f ms  %.0f GB/s  frac %.3f | step %.4f ms value %.2fM e2e %.2fM' % (r['avg_launch_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))
"; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/l_pytest.log
TSPN_SURVIVOR_PATH=0 timeout 300 python bench.py --no-cpu-baseline --steps 30 2> gpurun_out/l_s0.err | tee gpurun_out/l_s0.json | summ stored-rows
TSPN_SURVIVOR_PATH=1 timeout 300 python bench.py --no-cpu-baseline --steps 30 2> gpurun_out/l_s1.err | tee gpurun_out/l_s1.json | summ survivor
tail -3 gpurun_out/l_s1.err
TSPN_SURVIVOR_PATH=1 TSPN_SIDE_PRIORITY=0 timeout 300 python bench.py --no-cpu-baseline --steps 30 2> gpurun_out/l_s1p0.err | tee gpurun_out/l_s1p0.json | summ survivor-prio0
TSPN_SURVIVOR_PATH=1 timeout 300 python tools/trace_step.py --steps 2 > gpurun_out/l_trace.txt 2> gpurun_out/l_trace.err; tail -3 gpurun_out/l_trace.err
tail -24 gpurun_out/l_trace.txt
