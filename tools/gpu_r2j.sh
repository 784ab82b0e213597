# round 2, pass j: windowed kernel with the interior fast path - parity, bench, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_windowed.py -m gpu -q -x 2>&1 | tail -15
B="python bench.py --no-cpu-baseline --steps 20 --warmup 5 --geo-layout windowed"
for R in 0 8 24; do
$B --reserve-sms $R > gpurun_out/r2j_single_windowed_res$R.json 2> gpurun_out/r2j_single_windowed_res$R.err; tail -2 gpurun_out/r2j_single_windowed_res$R.err
done
python tools/trace_step.py --steps 1 --geo-layout windowed > gpurun_out/r2j_timeline_windowed.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pair_geo_windowed' -s 3 -c 1 -o gpurun_out/r2j_gw python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eager --geo-layout windowed > gpurun_out/r2j_gw_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2j_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("%-46s value %.1fM e2e %.1fM ms %.4f geo frac %.3f share %.3f launch %.4f alone %.4f (%.3f)" % (f, d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["frac"], r["share_of_step"], r["avg_launch_ms"], r["alone"]["avg_launch_ms"], r["alone"]["frac"]))
    except Exception as e: print(f, e)
PY
tail -20 gpurun_out/r2j_timeline_windowed.txt
