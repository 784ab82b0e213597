# round 2, pass n: batches of a step on two alternating compute streams (`value`), batch-size A/B
set -x
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-layout-extra"
$B --workload vidor_val --steps 5 --warmup 3 > gpurun_out/r2n_val.json 2> gpurun_out/r2n_val.err; tail -2 gpurun_out/r2n_val.err
$B --workload vidor_val --steps 5 --warmup 3 --serial-batches > gpurun_out/r2n_val_serial.json 2>/dev/null
$B --workload vidor_val --steps 5 --warmup 3 --geo-layout windowed > gpurun_out/r2n_val_windowed.json 2>/dev/null
$B --workload vidvrd_test --steps 10 --warmup 3 > gpurun_out/r2n_vrd.json 2> gpurun_out/r2n_vrd.err; tail -2 gpurun_out/r2n_vrd.err
$B --workload vidvrd_test --steps 10 --warmup 3 --serial-batches > gpurun_out/r2n_vrd_serial.json 2>/dev/null
$B --workload vidvrd_test --steps 10 --warmup 3 --max-videos 32 > gpurun_out/r2n_vrd_mv32.json 2>/dev/null
$B --workload vidvrd_test --steps 10 --warmup 3 --geo-layout windowed > gpurun_out/r2n_vrd_windowed.json 2>/dev/null
$B --steps 20 --warmup 5 > gpurun_out/r2n_single.json 2> gpurun_out/r2n_single.err; tail -2 gpurun_out/r2n_single.err
$B --steps 20 --warmup 5 --max-videos 8 > gpurun_out/r2n_single_mv8.json 2> gpurun_out/r2n_single_mv8.err; tail -2 gpurun_out/r2n_single_mv8.err
$B --steps 20 --warmup 5 --max-videos 4 > gpurun_out/r2n_single_mv4.json 2>/dev/null
$B --steps 20 --warmup 5 --max-videos 8 --geo-layout windowed > gpurun_out/r2n_single_mv8_windowed.json 2>/dev/null
$B --steps 20 --warmup 5 --max-videos 4 --geo-layout windowed > gpurun_out/r2n_single_mv4_windowed.json 2>/dev/null
$B --workload vidvrd_single --steps 20 --warmup 5 > gpurun_out/r2n_vrdsingle.json 2>/dev/null
$B --workload vidvrd_single --steps 20 --warmup 5 --max-videos 32 > gpurun_out/r2n_vrdsingle_mv32.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2n_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("%-46s value %.1fM e2e %.1fM ms %.4f geo frac %.3f share %.3f launch %.4f alone %.4f (%.3f) batches %d lanes %s" % (f, d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["frac"], r["share_of_step"], r["avg_launch_ms"], r["alone"]["avg_launch_ms"], r["alone"]["frac"], d["config"]["batches_per_step_per_gpu"], d["config"].get("batch_lanes")))
    except Exception as e: print(f, e)
PY
