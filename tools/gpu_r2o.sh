# round 2, pass o: full GPU test suite + default bench (with the windowed extra) + ragged workloads after the
# records-kernel changes + windowed kernel ncu / timeline
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > $O/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2o_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2o_bench_vidor_single.json 2> $O/r2o_bench_vidor_single.err; echo "bench rc=$?"; tail -2 $O/r2o_bench_vidor_single.err
for W in vidvrd_single vidvrd_test vidor_val; do
  timeout 900 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > $O/r2o_bench_$W.json 2> $O/r2o_bench_$W.err; echo "bench $W rc=$?"
done
timeout 300 python tools/trace_step.py --steps 1 --geo-layout windowed 2>/dev/null | grep -v arn > $O/r2o_timeline_vidor_single_windowed.txt
timeout 300 python tools/trace_step.py --steps 1 2>/dev/null | grep -v arn > $O/r2o_timeline_vidor_single.txt
timeout 300 python tools/trace_step.py --steps 1 --workload vidvrd_test --batches 2 2>/dev/null | grep -v arn > $O/r2o_timeline_vrd.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pair_geo_windowed' -s 3 -c 1 -o $O/r2o_gw python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-layout-extra --eager --geo-layout windowed > $O/r2o_gw_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2o_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("%-46s value %.1fM e2e %.1fM ms %.4f geo frac %.3f share %.3f launch %.4f alone %.4f (%.3f)" % (f, d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["frac"], r["share_of_step"], r["avg_launch_ms"], r["alone"]["avg_launch_ms"], r["alone"]["frac"]))
        if "windowed_layout" in d: print("    windowed_layout:", {k: d["windowed_layout"].get(k) for k in ("value","ms_per_step")})
        if "cpu_baseline" in d: print("    cpu_baseline:", d["cpu_baseline"])
    except Exception as e: print(f, e)
PY
head -30 $O/r2o_timeline_vrd.txt | cut -c1-120
