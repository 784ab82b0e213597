set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r_bench.json')); r=d['roofline']
print('value %.2fM e2e %.2fM step %.4f geo %.4f frac %.3f alone %.4f / %.3f launches %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['alone']['avg_launch_ms'], r['alone']['frac'], d['launches_per_step']))
"
timeout 300 python tools/trace_step.py --steps 2 > gpurun_out/r_timeline.txt 2> gpurun_out/r_timeline.err; tail -20 gpurun_out/r_timeline.txt
timeout 600 compute-sanitizer --tool initcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor.py -m gpu -x -q -k "survivor or kernel_shapes or phases or ragged" > gpurun_out/r_initcheck.log 2>&1; echo "initcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r_initcheck.log | head -5
