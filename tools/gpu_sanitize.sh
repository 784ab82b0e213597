# compute-sanitizer over the final round-1 code: memcheck on every GPU test file, initcheck / racecheck / synccheck on
# the tests of the kernels that changed last (persistent pair kernel, survivor rows, two-stream phases)
set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor.py tests/test_gpu_model.py tests/test_relations.py -m gpu -x -q \
  > gpurun_out/s3_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/s3_memcheck.log | head -10
for tool in initcheck racecheck synccheck; do
timeout 600 compute-sanitizer --tool $tool --print-limit 15 \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor.py -m gpu -x -q -k "survivor or kernel_shapes or phases or ragged" \
  > gpurun_out/s3_$tool.log 2>&1; echo "$tool rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/s3_$tool.log | head -5
grep -E "=========.*(Uninitialized|hazard|Error|error)" gpurun_out/s3_$tool.log | sort | uniq -c | sort -rn | head -12
done
