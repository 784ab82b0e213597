# compute-sanitizer over the parity tests (memcheck: out-of-bounds / misaligned accesses in every kernel)
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor.py tests/test_gpu_model.py tests/test_relations.py -m gpu -x -q \
  > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/sanitize_memcheck.log | head -20
tail -5 gpurun_out/sanitize_memcheck.log
