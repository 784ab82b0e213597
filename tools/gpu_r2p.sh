# round 2, pass p: windowed kernel with the deferred queue read / linear row advance; stored-rows path on small videos
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_windowed.py tests/test_gpu_ragged.py -m gpu -q -x 2>&1 | tail -3
B="python bench.py --no-cpu-baseline --no-layout-extra"
$B --steps 20 --warmup 5 --geo-layout windowed > $O/r2p_single_windowed.json 2> $O/r2p_single_windowed.err; tail -2 $O/r2p_single_windowed.err
$B --workload vidor_val --steps 5 --warmup 3 --geo-layout windowed > $O/r2p_val_windowed.json 2>/dev/null
$B --workload vidvrd_test --steps 10 --warmup 3 --geo-layout windowed > $O/r2p_vrd_windowed.json 2>/dev/null
TSPN_SURVIVOR_PATH=0 $B --workload vidvrd_test --steps 10 --warmup 3 > $O/r2p_vrd_storedrows.json 2> $O/r2p_vrd_storedrows.err; tail -2 $O/r2p_vrd_storedrows.err
TSPN_SURVIVOR_PATH=0 $B --workload vidvrd_single --steps 20 --warmup 5 > $O/r2p_vrdsingle_storedrows.json 2>/dev/null
TSPN_SURVIVOR_PATH=0 $B --workload vidor_val --steps 5 --warmup 3 > $O/r2p_val_storedrows.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2p_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("%-46s value %.1fM e2e %.1fM ms %.4f geo frac %.3f share %.3f launch %.4f alone %.4f (%.3f)" % (f, d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["frac"], r["share_of_step"], r["avg_launch_ms"], r["alone"]["avg_launch_ms"], r["alone"]["frac"]))
    except Exception as e: print(f, e)
PY
