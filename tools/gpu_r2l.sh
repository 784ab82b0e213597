# round 2, pass l: windowed kernel - pairs per work unit A/B
set -x
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --steps 20 --warmup 5 --geo-layout windowed --reserve-sms 8"
run() { # tag, nvcc extra
  touch temporal-span-proposal-network-vidvrd_b200/csrc/geo_windowed.cu
  TSPN_NVCC_EXTRA="$2" python -m tspn_b200.build > gpurun_out/r2l_build_$1.log 2>&1; tail -1 gpurun_out/r2l_build_$1.log
  $B > gpurun_out/r2l_single_$1.json 2> gpurun_out/r2l_single_$1.err; tail -2 gpurun_out/r2l_single_$1.err
}
run u1 "-DTSPN_GW_UNIT=1"
run u2 "-DTSPN_GW_UNIT=2"
run u4 "-DTSPN_GW_UNIT=4"
run u2r96 "-DTSPN_GW_UNIT=2 -DTSPN_GW_MAXNREG=96"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2l_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("%-46s value %.1fM e2e %.1fM ms %.4f geo frac %.3f share %.3f launch %.4f alone %.4f (%.3f)" % (f, d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["frac"], r["share_of_step"], r["avg_launch_ms"], r["alone"]["avg_launch_ms"], r["alone"]["frac"]))
    except Exception as e: print(f, e)
PY
