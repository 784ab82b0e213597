set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/c_pytest.log
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1: geo %.4f ms  %.0f GB/s  frac %.3f | step %.4f ms value %.2fM e2e %.2fM' % (r['avg_launch_ms'], r['achieved'], r['frac'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))
"; }
timeout 300 python bench.py --no-cpu-baseline --steps 40 2> gpurun_out/c.err | tee gpurun_out/c_bench_cs2.json | summ cs2
timeout 300 python bench.py --no-cpu-baseline --steps 40 --compute-streams 1 2>> gpurun_out/c.err | tee gpurun_out/c_bench_cs1.json | summ cs1
timeout 300 python bench.py --no-cpu-baseline --steps 40 --depth 4 2>> gpurun_out/c.err | tee gpurun_out/c_bench_cs2_d4.json | summ cs2_depth4
python - <<'PY'
import torch, time
a = torch.empty(49489408, dtype=torch.uint8).pin_memory(); b = torch.empty_like(a, device='cuda')
c = torch.empty(18518080, dtype=torch.uint8, device='cuda'); d = torch.empty(18518080, dtype=torch.uint8).pin_memory()
for _ in range(3): b.copy_(a, non_blocking=True); d.copy_(c, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): b.copy_(a, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D 49.5MB: %.3f ms  %.1f GB/s" % (e0.elapsed_time(e1)/20, 49489408/(e0.elapsed_time(e1)/20*1e6)))
e0.record()
for _ in range(20): d.copy_(c, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("D2H 18.5MB: %.3f ms  %.1f GB/s" % (e0.elapsed_time(e1)/20, 18518080/(e0.elapsed_time(e1)/20*1e6)))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eager > gpurun_out/c_b_ncu.log 2>&1
tail -5 gpurun_out/c.err
