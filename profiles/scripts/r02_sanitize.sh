# compute-sanitizer passes over the paths added in round 2 (small cases; memcheck + racecheck on the shared-memory kernels)
set -x
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest -x -q \
  "tests/test_gpu_window.py::test_windowed_decode_equals_the_oracle[801-400]" \
  "tests/test_gpu_pipelining.py::test_multi_chunk_calls_alternate_slots[True]" \
  "tests/test_gpu_simd_sat.py::test_saturating_flavour_streaming_rows" \
  "tests/test_gpu_tag_stress.py::test_tagged_decision_row_kernel_on_tie_storms[Voyager-HARD8]" 2>&1 | tail -8
echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest -x -q \
  "tests/test_gpu_history_k15.py::test_k15_history_is_the_default_for_soft16_batches" \
  "tests/test_gpu_history_k9.py" -k "default or residue" 2>&1 | tail -8
echo "racecheck rc=$?"
