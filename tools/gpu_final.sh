# Round-2 evidence run on ONE B200: tests, every bench workload (each line carries the windowed-layout step beside
# the dense one), reference arm, timelines, launch list, ncu captures.  Writes gpurun_out/r2_*.
# usage: bash tools/gpu_final.sh [full]      (full: also the sustained run, the contraction-kernel micro-benchmarks,
#                                             the single-video latency and the side-branch ncu captures)
set -x
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1700 python -m pytest tests -m gpu -q > $O/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_vidor_single.json 2> $O/r2_bench_vidor_single.err; echo "bench rc=$?"; tail -2 $O/r2_bench_vidor_single.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
for W in vidvrd_single vidvrd_test vidor_val stress; do
  timeout 900 python bench.py --workload $W --steps 5 --warmup 3 > $O/r2_bench_$W.json 2> $O/r2_bench_$W.err; echo "bench $W rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --geo-layout windowed > $O/r2_bench_vidor_single_windowed.json 2>/dev/null
timeout 300 python tools/trace_step.py --steps 1 --geo-layout windowed 2>/dev/null | grep -v arn > $O/r2_timeline_vidor_single_windowed.txt
timeout 300 python tools/trace_step.py --steps 2 2>/dev/null | grep -v arn > $O/r2_timeline_vidor_single.txt
timeout 300 python tools/trace_step.py --steps 1 --workload vidor_val --batches 3 2>/dev/null | grep -v arn > $O/r2_timeline_vidor_val.txt
timeout 300 python tools/trace_step.py --steps 1 --workload vidvrd_test --batches 2 2>/dev/null | grep -v arn > $O/r2_timeline_vidvrd_test.txt
# launch list of the bench command (eager launches: every kernel is a launch), full captures of the two all-pairs kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-layout-extra --eager > $O/r2_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_geo_kernel' -s 2 -c 1 -o $O/r2_prof_pair python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-layout-extra --eager > $O/r2_prof.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pair_geo_windowed' -s 3 -c 1 -o $O/r2_prof_windowed python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-layout-extra --eager --geo-layout windowed > /dev/null 2>&1
if [ "$1" = "full" ]; then
for W in baseline_yaml; do timeout 900 python bench.py --workload $W --steps 5 --warmup 3 > $O/r2_bench_$W.json 2> $O/r2_bench_$W.err; done
timeout 600 python bench.py --workload baseline_yaml --precision tensor --steps 5 --warmup 3 > $O/r2_bench_baseline_yaml_tensor.json 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-layout-extra --span-proposals 0 > $O/r2_bench_vidor_single_no_nms.json 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-layout-extra --relationness tensor > $O/r2_bench_vidor_single_tc_relationness.json 2>/dev/null
timeout 600 python bench.py --steps 2000 --warmup 20 --no-cpu-baseline --no-layout-extra > $O/r2_bench_vidor_single_sustained.json 2>/dev/null
timeout 300 python tools/trace_step.py --steps 1 --relationness tensor 2>/dev/null | grep -v arn > $O/r2_timeline_vidor_single_tc.txt
timeout 300 python tools/trace_step.py --steps 1 --workload stress 2>/dev/null | grep -v arn > $O/r2_timeline_stress.txt
timeout 600 python tools/latency_basemodel.py --calls 200 > $O/r2_latency_basemodel.jsonl 2> $O/r2_latency.err; tail -2 $O/r2_latency.err
timeout 600 python tools/bench_predicate.py > $O/r2_predicate.jsonl 2> $O/r2_predicate.err; cat $O/r2_predicate.jsonl
timeout 600 python tools/bench_span_head.py 256 1024 300 > $O/r2_span_head.jsonl 2> $O/r2_span_head.err
timeout 600 python tools/bench_span_head.py 1024 1024 2000 >> $O/r2_span_head.jsonl 2>> $O/r2_span_head.err; cat $O/r2_span_head.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'span_select|survivor_rows|scores_topk|ppn_embed|pair_top_predicates|predicate_tc' -s 20 -c 8 -o $O/r2_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-layout-extra --eager > $O/r2_prof_side.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'predicate_tc' -c 2 -o $O/r2_prof_predicate python tools/bench_predicate.py > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'span_head_tc' -c 1 -o $O/r2_prof_span_head python tools/bench_span_head.py 256 1024 300 > /dev/null 2>&1
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        d=json.load(open(f))
        if d.get("impl") == "reference": print(f, d["value"], d.get("cpu_baseline", {}).get("sample", "")[:80]); continue
        r=d["roofline"]
        print("%-52s value %.1fM e2e %.1fM ms %.4f frac %.3f alone %.3f" % (f, d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["frac"], r.get("alone",{}).get("frac",0)), "| windowed:", round(d.get("windowed_layout",{}).get("value") or 0)/1e6)
    except Exception as e: print(f, e)
PY
