# final checks of the round: GPU tests, memcheck over every GPU test file but the full-size one, bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor.py tests/test_gpu_model.py tests/test_relations.py -m gpu -x -q \
  > gpurun_out/f_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/f_memcheck.log | head -6
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor.py -m gpu -x -q -k "survivor or fused or reserved or phases" > gpurun_out/f_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/f_racecheck.log | head -4
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline 2> gpurun_out/f_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.2fM e2e %.2fM step %.4f geo %.4f frac %.3f alone %.4f / %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['alone']['avg_launch_ms'], r['alone']['frac']))
"
