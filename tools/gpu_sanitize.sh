# compute-sanitizer over the round-2 code: memcheck on the GPU test files of the kernels that changed (sentinel
# totals / capacity grids, chunk slots, streaming survivor kernel, span selection, tcgen05 relationness, span-packed
# boxes), racecheck / synccheck / initcheck on the shared-memory heavy ones.
set -x
mkdir -p gpurun_out
if [ "$1" != "b" ] && [ "$1" != "c" ]; then
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_ragged.py tests/test_gpu_relationness_tc.py tests/test_gpu_parity.py tests/test_gpu_tensor.py tests/test_gpu_model.py -m gpu -x -q \
  > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/r2_memcheck.log | head -10
for tool in racecheck synccheck initcheck; do
timeout 900 compute-sanitizer --tool $tool --print-limit 15 \
  python -m pytest tests/test_gpu_ragged.py tests/test_gpu_relationness_tc.py tests/test_gpu_tensor.py -m gpu -x -q -k "span_select or capacity or survivor or graph or scores_within or fused_tc" \
  > gpurun_out/r2_$tool.log 2>&1; echo "$tool rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_$tool.log | head -5
grep -E "=========.*(Uninitialized|hazard|Error|error)" gpurun_out/r2_$tool.log | sort | uniq -c | sort -rn | head -12
done
fi
if [ "$1" != "c" ]; then
# ---- second half of the round: the windowed layout's kernels (per-warp cp.async double buffers, the pair queue, the
# offsets scan), the REDUX records kernel and the 2-D predicate reduce ------------------------------------------------
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_windowed.py tests/test_gpu_parity.py tests/test_gpu_tensor.py -m gpu -x -q -k "windowed or records or postprocess or predicate or triplet" \
  > gpurun_out/r2b_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/r2b_memcheck.log | head -10
for tool in racecheck synccheck initcheck; do
timeout 900 compute-sanitizer --tool $tool --print-limit 15 \
  python -m pytest tests/test_gpu_windowed.py -m gpu -x -q -k "rows_equal_dense or degenerate or capacity_graph" \
  > gpurun_out/r2b_$tool.log 2>&1; echo "$tool rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2b_$tool.log | head -5
grep -E "=========.*(Uninitialized|hazard|Error|error)" gpurun_out/r2b_$tool.log | sort | uniq -c | sort -rn | head -12
done
fi
# ---- last part of the round (bash tools/gpu_sanitize.sh c): the delta-coded box expansion (block scan through shared
# memory) and the tensor-core span head (staged epilogue slice, named barrier, CTA-pair form: cluster barriers, remote
# mbarrier arrives, multicast commits) -----------------------------------------------------------------------------------
SEL="delta_coded or compact_transport or span_head_tensor"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_model.py tests/test_gpu_tensor.py -m gpu -x -q -k "$SEL" \
  > gpurun_out/r2c_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/r2c_memcheck.log | head -10
for tool in racecheck synccheck; do
timeout 600 compute-sanitizer --tool $tool --print-limit 15 \
  python -m pytest tests/test_gpu_model.py tests/test_gpu_tensor.py -m gpu -x -q -k "$SEL" \
  > gpurun_out/r2c_$tool.log 2>&1; echo "$tool rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2c_$tool.log | head -5
grep -E "=========.*(hazard|Error|error)" gpurun_out/r2c_$tool.log | sort | uniq -c | sort -rn | head -12
done
