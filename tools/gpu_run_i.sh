set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/i_pytest.log
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1: geo %.4f ms  %.0f GB/s  frac %.3f | step %.4f ms value %.2fM e2e %.2fM' % (r['avg_launch_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))
"; }
timeout 300 python bench.py --no-cpu-baseline --steps 40 2> gpurun_out/i.err | tee gpurun_out/i_bench.json | summ decomposed
timeout 300 python bench.py --no-cpu-baseline --steps 40 2>> gpurun_out/i.err | summ decomposed_again
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eager > gpurun_out/i_b_ncu.log 2>&1
tail -5 gpurun_out/i.err
