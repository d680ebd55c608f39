#!/bin/bash
# compute-sanitizer pass over a representative subset of the GPU tests (memcheck: out-of-bounds / misaligned accesses in
# every kernel of the library; racecheck on the search kernels' and the loss kernel's shared-memory hand-offs).  Summaries go to gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
mkdir -p gpurun_out
SEL="tests/test_search_gpu.py::test_integer_data_exact tests/test_search_gpu.py::test_small_duplicate_cluster_resolved_by_wide_level tests/test_search_gpu.py::test_near_duplicate_cluster_is_exact tests/test_search_gpu.py::test_three_phase_sharded_search_prunes_and_stays_exact tests/test_search_gpu.py::test_pair_scan_and_single_cta_scan_agree[9000-64-257-10] tests/test_loss_gpu.py tests/test_encoder_gpu.py::test_bert_small_matches_reference_golden tests/test_encoder_gpu.py::test_t5_small_matches_reference_golden"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest $SEL -m gpu -x -q > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_memcheck.log
tail -5 gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 20 python -m pytest "tests/test_search_gpu.py::test_integer_data_exact" "tests/test_search_gpu.py::test_massive_ties" "tests/test_search_gpu.py::test_pair_scan_and_single_cta_scan_agree[9000-64-257-10]" "tests/test_loss_gpu.py::test_gradients_are_run_to_run_identical" -m gpu -x -q > gpurun_out/r02_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_racecheck.log
tail -5 gpurun_out/r02_racecheck.log
