# compute-sanitizer over the double-buffered K=15 kernels (round 2, second half): racecheck on the speculative path, the rollback /
# replay path (thresholds that renormalise in every group or every few groups) and the decision-row kernel (batch + streaming);
# memcheck on the same
set -x
export PYTHONUNBUFFERED=1
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest -x -q \
  "tests/test_gpu_history_k15.py::test_k15_history_is_the_default_for_soft16_batches" \
  "tests/test_gpu_history_k15.py::test_k15_history_renormalisation_rollback[low]" \
  "tests/test_gpu_history_k15.py::test_k15_history_renormalisation_rollback[one]" \
  "tests/test_gpu_history_k15.py::test_k15_history_segmented_traceback_and_decision_row_kernels_agree" 2>&1 | tail -8
echo "racecheck rc=$?"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest -x -q \
  "tests/test_gpu_history_k15.py::test_k15_history_renormalisation_rollback[low]" \
  "tests/test_gpu_history_k15.py::test_k15_history_segmented_traceback_and_decision_row_kernels_agree" \
  "tests/test_gpu_window.py::test_windowed_decode_equals_the_oracle[801-400]" 2>&1 | tail -8
echo "memcheck rc=$?"
