set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_survivor.py 16 2>&1 | tail -4
for v in 8 32; do
TSPN_SURVIVOR_PATH=1 timeout 300 python tools/trace_step.py --steps 1 --videos $v > gpurun_out/q_trace_$v.txt 2> gpurun_out/q_trace_$v.err
grep -E "== step|pair_geo_kernel|survivor_rows|video_top|topk_kernel" gpurun_out/q_trace_$v.txt | tail -6
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'survivor_rows' -c 1 -o gpurun_out/q_surv python tools/bench_survivor.py 16 > gpurun_out/q_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/q_surv.ncu-rep --page raw --csv > gpurun_out/q_surv_raw.csv 2>/dev/null
ls -la gpurun_out | grep q_
