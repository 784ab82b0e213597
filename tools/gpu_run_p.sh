set -x
mkdir -p gpurun_out
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1: geo %.4f ms  %.0f GB/s  frac %.3f | step %.4f ms value %.2fM e2e %.2fM' % (r['avg_launch_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))
"; }
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 30 2> gpurun_out/p_$tag.err | tee gpurun_out/p_$tag.json | summ $tag || tail -3 gpurun_out/p_$tag.err; }
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/p_pytest.log
run surv0 TSPN_SURVIVOR_PATH=0
run surv1 TSPN_SURVIVOR_PATH=1
TSPN_SURVIVOR_PATH=1 timeout 300 python tools/trace_step.py --steps 2 > gpurun_out/p_trace.txt 2> gpurun_out/p_trace.err
tail -18 gpurun_out/p_trace.txt
