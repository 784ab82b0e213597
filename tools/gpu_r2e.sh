set -x
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline"
run() { # tag, nvcc extra
  TSPN_NVCC_EXTRA="$2" python -m tspn_b200.build --force > gpurun_out/r2e_build_$1.log 2>&1; tail -1 gpurun_out/r2e_build_$1.log
  $B --steps 20 --warmup 5 > gpurun_out/r2e_single_$1.json 2> gpurun_out/r2e_single_$1.err; tail -2 gpurun_out/r2e_single_$1.err
  $B --workload vidor_val --steps 5 --warmup 3 > gpurun_out/r2e_val_$1.json 2>/dev/null
  python tools/trace_step.py --steps 1 > gpurun_out/r2e_timeline_$1.txt 2>&1
}
run g104s5 "-DTSPN_GEO_MAXNREG=104 -DTSPN_SV_MIN_CTAS=5"
run g88s6 "-DTSPN_GEO_MAXNREG=88 -DTSPN_SV_MIN_CTAS=6"
run g96s6 "-DTSPN_GEO_MAXNREG=96 -DTSPN_SV_MIN_CTAS=6"
run g88s5 "-DTSPN_GEO_MAXNREG=88 -DTSPN_SV_MIN_CTAS=5"
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_tensor.py tests/test_headline_golden.py -m gpu -q 2>&1 | tail -5
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2e_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("%-40s value %.1fM e2e %.1fM ms %.4f geo frac %.3f share %.3f launch %.4f alone %.4f" % (f, d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["frac"], r["share_of_step"], r["avg_launch_ms"], r["alone"]["avg_launch_ms"]))
    except Exception as e: print(f, e)
PY
