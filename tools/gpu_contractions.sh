# The two tcgen05 contraction kernels at the sizes that load them: parity tests, CUDA-event times, ncu captures.
# usage: bash tools/gpu_contractions.sh      (writes gpurun_out/r2_span_head.jsonl, r2_predicate.jsonl, r2_prof_*.ncu-rep)
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_tensor.py -m gpu -q -x > $O/r2_pytest_tensor.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_pytest_tensor.log
timeout 200 python tools/bench_span_head.py 256 1024 300 > $O/r2_span_head.jsonl 2> $O/r2_span_head.err
timeout 200 python tools/bench_span_head.py 1024 1024 2000 >> $O/r2_span_head.jsonl 2>> $O/r2_span_head.err
timeout 300 python tools/bench_predicate.py > $O/r2_predicate.jsonl 2> $O/r2_predicate.err
cat $O/r2_span_head.jsonl $O/r2_predicate.jsonl | cut -c1-330
timeout 600 ncu --set full --clock-control none -k regex:'span_head_tc' -s 6 -c 2 -o $O/r2_prof_span_head -f python tools/bench_span_head.py 256 1024 300 > /dev/null 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'span_' -c 40 --csv --log-file $O/r2_launches_span_head.csv python tools/bench_span_head.py 256 1024 300 > /dev/null 2>&1
