set -x
mkdir -p gpurun_out
for tool in initcheck racecheck synccheck; do
timeout 900 compute-sanitizer --tool $tool --print-limit 15 \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_model.py -m gpu -x -q \
  > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | head -5
grep -E "=========.*(Uninitialized|hazard|Error|error)" gpurun_out/sanitize_$tool.log | sort | uniq -c | sort -rn | head -12
done
