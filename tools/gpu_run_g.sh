set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/g_pytest.log
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1: geo %.4f ms  %.0f GB/s  frac %.3f | step %.4f ms value %.2fM e2e %.2fM h2d %.1fMB' % (r['avg_launch_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['h2d_bytes_per_step']/1e6))
"; }
timeout 300 python bench.py --no-cpu-baseline --steps 40 2> gpurun_out/g.err | tee gpurun_out/g_bench_compact.json | summ compact
timeout 300 python bench.py --no-cpu-baseline --steps 40 --fp32-transport 2>> gpurun_out/g.err | tee gpurun_out/g_bench_fp32.json | summ fp32
timeout 300 python bench.py --no-cpu-baseline --steps 40 --compute-streams 2 2>> gpurun_out/g.err | summ compact_cs2
timeout 300 python bench.py --no-cpu-baseline --steps 40 --depth 2 2>> gpurun_out/g.err | summ compact_depth2
tail -5 gpurun_out/g.err
