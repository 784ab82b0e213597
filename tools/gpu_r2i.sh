# round 2, pass i: windowed warp-per-pair kernel - ncu capture, occupancy variants
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pair_geo_windowed' -s 3 -c 1 -o gpurun_out/r2i_gw python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eager --geo-layout windowed > gpurun_out/r2i_gw_ncu.log 2>&1; echo "ncu rc=$?"
B="python bench.py --no-cpu-baseline --steps 20 --warmup 5 --geo-layout windowed --reserve-sms 0"
run() { # tag, nvcc extra
  touch temporal-span-proposal-network-vidvrd_b200/csrc/geo_windowed.cu
  TSPN_NVCC_EXTRA="$2" python -m tspn_b200.build > gpurun_out/r2i_build_$1.log 2>&1; tail -1 gpurun_out/r2i_build_$1.log
  $B > gpurun_out/r2i_single_$1.json 2> gpurun_out/r2i_single_$1.err; tail -2 gpurun_out/r2i_single_$1.err
}
run c3r80 "-DTSPN_GW_CTAS=3 -DTSPN_GW_MAXNREG=80"
run c4r64 "-DTSPN_GW_CTAS=4 -DTSPN_GW_MAXNREG=64"
run c2r128 "-DTSPN_GW_CTAS=2 -DTSPN_GW_MAXNREG=128"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2i_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("%-46s value %.1fM e2e %.1fM ms %.4f geo frac %.3f share %.3f launch %.4f alone %.4f (%.3f)" % (f, d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["frac"], r["share_of_step"], r["avg_launch_ms"], r["alone"]["avg_launch_ms"], r["alone"]["frac"]))
    except Exception as e: print(f, e)
PY
