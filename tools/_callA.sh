set -x
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r2s_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2s_single_delta.json 2> $O/r2s_single_delta.err; echo rc=$?; tail -2 $O/r2s_single_delta.err
TSPN_BOX_DELTA=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-layout-extra > $O/r2s_single_raw.json 2>/dev/null; echo rc=$?
timeout 300 python tools/trace_step.py --steps 2 --workload vidvrd_test --batches 2 2>/dev/null | grep -v arn > $O/r2s_timeline_vrd.txt
python - <<'PY'
import json
for f in ("r2s_single_delta","r2s_single_raw"):
    d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"]/1e6, d["e2e"], d["ms_per_step"], d["roofline"]["frac"])
PY
