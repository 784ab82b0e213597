O=gpurun_out
rm -f $O/r2w_span_head.jsonl
export TSPN_SPAN_HEAD_ONE_CTA=1
for D in 0 3 7 8 11 15 4 12; do
  echo "one_cta=1 dbg=$D" >> $O/r2w_span_head.jsonl
  TSPN_SPAN_HEAD_DEBUG=$D timeout 120 python tools/bench_span_head.py 256 1024 300 >> $O/r2w_span_head.jsonl 2>> $O/r2w_span_head.err
done
cut -c1-130 $O/r2w_span_head.jsonl
