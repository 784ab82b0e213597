set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2> gpurun_out/s_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.2fM e2e %.2fM step %.4f geo %.4f frac %.3f alone %.4f / %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['alone']['avg_launch_ms'], r['alone']['frac']))
"
done
timeout 300 python tools/trace_step.py --steps 2 > gpurun_out/s_timeline.txt 2> gpurun_out/s_timeline.err; tail -20 gpurun_out/s_timeline.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --eager > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/s_launches.csv')) if len(r)>10 and r[0].isdigit()]
seen={}
for r in rows[-20:]:
    print(r[4][:60], r[-1])
PY
