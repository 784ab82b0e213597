# round 2, pass h: warp-per-pair windowed kernel - parity (both kernels), bench, timeline
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_windowed.py -m gpu -q -x 2>&1 | tail -15
TSPN_GEO_WIN_RING=1 timeout 900 python -m pytest tests/test_gpu_windowed.py -m gpu -q -x 2>&1 | tail -3
B="python bench.py --no-cpu-baseline --steps 20 --warmup 5 --geo-layout windowed"
for R in 8 0 16 32 64; do
$B --reserve-sms $R > gpurun_out/r2h_single_windowed_res$R.json 2> gpurun_out/r2h_single_windowed_res$R.err; tail -2 gpurun_out/r2h_single_windowed_res$R.err
done
python bench.py --no-cpu-baseline --workload vidor_val --steps 5 --warmup 3 --geo-layout windowed > gpurun_out/r2h_val_windowed.json 2> gpurun_out/r2h_val_windowed.err; tail -2 gpurun_out/r2h_val_windowed.err
python bench.py --no-cpu-baseline --workload stress --steps 10 --warmup 3 --geo-layout windowed > gpurun_out/r2h_stress_windowed.json 2> gpurun_out/r2h_stress_windowed.err; tail -2 gpurun_out/r2h_stress_windowed.err
python bench.py --no-cpu-baseline --workload vidvrd_test --steps 10 --warmup 3 --geo-layout windowed > gpurun_out/r2h_vrd_windowed.json 2> gpurun_out/r2h_vrd_windowed.err; tail -2 gpurun_out/r2h_vrd_windowed.err
python tools/trace_step.py --steps 1 --geo-layout windowed > gpurun_out/r2h_timeline_windowed.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2h_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("%-46s value %.1fM e2e %.1fM ms %.4f geo frac %.3f share %.3f launch %.4f alone %.4f (%.3f)" % (f, d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["frac"], r["share_of_step"], r["avg_launch_ms"], r["alone"]["avg_launch_ms"], r["alone"]["frac"]))
    except Exception as e: print(f, e)
PY
tail -22 gpurun_out/r2h_timeline_windowed.txt
