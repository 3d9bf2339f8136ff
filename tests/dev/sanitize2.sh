#!/bin/bash
# compute-sanitizer over the decode-side robustness tests of the PRODUCT library (memcheck + racecheck) and memcheck over the
# kernels added late in round 2 (one-warp decode CTAs run in every test; gnib / gtagg through the variant suite, A/B library).
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
AB=$PWD/lz4_flex_b200/liblz4b200_ab.so
rm -f gpurun_out/sanitize_summary.txt
T1="tests/test_gpu_block.py::test_garbage_matches_oracle tests/test_gpu_block.py::test_no_output_leak"
T2="tests/test_gpu_dict.py tests/test_gpu_frame.py::test_linked_frame_errors_match_oracle tests/test_gpu_frame.py::test_linked_frames_decode_on_gpu"
for tool in memcheck racecheck; do
  for grp in 1 2; do
    eval tests=\$T$grp
    timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 0 python -m pytest $tests -q -p no:cacheprovider -x -m gpu > gpurun_out/sanitize_${tool}_$grp.log 2>&1
    { echo "== compute-sanitizer --tool $tool :: $tests"; grep -E " passed| failed|ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" gpurun_out/sanitize_${tool}_$grp.log | sort | uniq -c | sort -rn | head -12; } >> gpurun_out/sanitize_summary.txt
  done
done
LZ4B200_SO_OVERRIDE=$AB timeout 1200 $CS --tool memcheck --print-limit 20 --error-exitcode 0 python -m pytest tests/variants_impl.py -q -p no:cacheprovider -x -m gpu -k "(nib or tag8 or tagg) and (compress_all_modes or many_blocks)" > gpurun_out/sanitize_memcheck_tags.log 2>&1
{ echo "== compute-sanitizer --tool memcheck :: variants_impl.py -k '(nib or tag8 or tagg) and (compress_all_modes or many_blocks)' (A/B library)"; grep -E " passed| failed|ERROR SUMMARY|Error:" gpurun_out/sanitize_memcheck_tags.log | sort | uniq -c | sort -rn | head -8; } >> gpurun_out/sanitize_summary.txt
cat gpurun_out/sanitize_summary.txt
