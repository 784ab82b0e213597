O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor.py -m gpu -q -x > $O/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2y_pytest.log
rm -f $O/r2y_span_head.jsonl
for ONE in 0 1; do for D in 0 1 2 3; do
  if [ $ONE = 1 ]; then export TSPN_SPAN_HEAD_ONE_CTA=1; else unset TSPN_SPAN_HEAD_ONE_CTA; fi
  echo "one_cta=$ONE dbg=$D" >> $O/r2y_span_head.jsonl
  TSPN_SPAN_HEAD_DEBUG=$D timeout 120 python tools/bench_span_head.py 256 1024 300 >> $O/r2y_span_head.jsonl 2>> $O/r2y_span_head.err
done; done
for ONE in 0 1; do
  if [ $ONE = 1 ]; then export TSPN_SPAN_HEAD_ONE_CTA=1; else unset TSPN_SPAN_HEAD_ONE_CTA; fi
  echo "one_cta=$ONE big" >> $O/r2y_span_head.jsonl
  timeout 120 python tools/bench_span_head.py 1024 1024 2000 >> $O/r2y_span_head.jsonl 2>> $O/r2y_span_head.err
done
unset TSPN_SPAN_HEAD_ONE_CTA
timeout 300 python tools/bench_predicate.py > $O/r2y_predicate.jsonl 2> $O/r2y_predicate.err
cut -c1-130 $O/r2y_span_head.jsonl; cut -c1-300 $O/r2y_predicate.jsonl
