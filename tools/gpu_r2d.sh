set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -15
B="python bench.py --no-cpu-baseline"
$B --steps 20 --warmup 5 > gpurun_out/r2d_single.json 2> gpurun_out/r2d_single.err; tail -2 gpurun_out/r2d_single.err
$B --workload vidor_val --steps 5 --warmup 3 > gpurun_out/r2d_val.json 2> gpurun_out/r2d_val.err; tail -2 gpurun_out/r2d_val.err
for R in 16 24 40; do $B --workload vidor_val --steps 5 --warmup 3 --reserve-sms $R > gpurun_out/r2d_val_res$R.json 2>/dev/null; done
$B --workload vidvrd_test --steps 5 --warmup 3 > gpurun_out/r2d_vrd.json 2>/dev/null
python tools/trace_step.py --steps 1 > gpurun_out/r2d_timeline.txt 2>&1
python tools/trace_step.py --steps 1 --workload vidor_val --batches 3 > gpurun_out/r2d_timeline_val.txt 2>&1
bash tools/ncu_hot.sh r2d
# the pair kernel at 80 registers: two side-branch CTAs of 96 registers fit beside it
TSPN_NVCC_EXTRA="-DTSPN_GEO_MAXNREG=80" python -m tspn_b200.build --force > gpurun_out/r2d_build80.log 2>&1; tail -2 gpurun_out/r2d_build80.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor.py -m gpu -q 2>&1 | tail -3
$B --steps 20 --warmup 5 > gpurun_out/r2d_single_r80.json 2> gpurun_out/r2d_single_r80.err; tail -2 gpurun_out/r2d_single_r80.err
$B --workload vidor_val --steps 5 --warmup 3 > gpurun_out/r2d_val_r80.json 2>/dev/null
$B --workload stress --steps 5 --warmup 3 > gpurun_out/r2d_stress_r80.json 2>/dev/null
python tools/trace_step.py --steps 1 > gpurun_out/r2d_timeline_r80.txt 2>&1
python tools/trace_step.py --steps 1 --workload vidor_val --batches 3 > gpurun_out/r2d_timeline_val_r80.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2d_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("%-40s value %.1fM e2e %.1fM ms %.4f geo frac %.3f share %.3f launch %.4f" % (f, d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["frac"], r["share_of_step"], r["avg_launch_ms"]))
    except Exception as e: print(f, e)
PY
