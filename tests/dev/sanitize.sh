#!/bin/bash
# compute-sanitizer over the decode-side robustness tests (the reference's analogue: MSan on fuzz_decomp_corrupt_block,
# .github/workflows/rust.yml:52-60) and over the kernels that hand data between warps (mbarrier queues, TMA ring, lane-group
# matchers under sub-warp masks, the linked decoder's spin-wait).  Summaries land in gpurun_out/sanitize_summary.txt; copy
# them to profiles/.  The variant suite runs against the A/B library so every kernel family is covered.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
AB=$PWD/lz4_flex_b200/liblz4b200_ab.so
T1="tests/test_gpu_block.py::test_garbage_matches_oracle tests/test_gpu_block.py::test_no_output_leak"
T2="tests/test_gpu_dict.py tests/test_gpu_frame.py::test_linked_frame_errors_match_oracle tests/test_gpu_frame.py::test_linked_frames_decode_on_gpu"
T3="tests/variants_impl.py::test_decode_errors_match_oracle tests/variants_impl.py::test_overlapping_periods"
T4="tests/variants_impl.py::test_compress_all_modes_and_roundtrip"
for tool in memcheck racecheck; do
  for grp in 1 2 3 4; do
    eval tests=\$T$grp
    so=""; if [ $grp -ge 3 ]; then so=$AB; fi
    LZ4B200_SO_OVERRIDE=$so timeout 1200 $CS --tool $tool --print-limit 20 --error-exitcode 0 python -m pytest $tests -q -p no:cacheprovider -x -m gpu > gpurun_out/sanitize_${tool}_$grp.log 2>&1
    { echo "== compute-sanitizer --tool $tool :: $tests"; grep -E " passed| failed|ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" gpurun_out/sanitize_${tool}_$grp.log | sort | uniq -c | sort -rn | head -12; } >> gpurun_out/sanitize_summary.txt
  done
done
cat gpurun_out/sanitize_summary.txt
