O=gpurun_out
rm -f $O/r2x_span_head.jsonl
export TSPN_SPAN_HEAD_ONE_CTA=1
for S in 2 3 4; do for D in 0 7; do
  echo "one_cta=1 stages=$S dbg=$D" >> $O/r2x_span_head.jsonl
  TSPN_SPAN_HEAD_STAGES=$S TSPN_SPAN_HEAD_DEBUG=$D timeout 120 python tools/bench_span_head.py 256 1024 300 >> $O/r2x_span_head.jsonl 2>> $O/r2x_span_head.err
done; done
cut -c1-130 $O/r2x_span_head.jsonl
