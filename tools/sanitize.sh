#!/bin/bash
# compute-sanitizer passes over a representative subset of the GPU tests (run under gpurun from the repo root):
#   bash tools/sanitize.sh            -> gpurun_out/sanitize_{memcheck,racecheck}.log
# memcheck: out-of-bounds / misaligned accesses of every kernel of the path; racecheck: shared-memory hazards of the
# CTA-cooperative kernels (FAST cells, oct-tree, describe, window query, stereo, distinctive descriptors).
OUT=gpurun_out
mkdir -p $OUT
SEL='golden_cases and (c1_seed1000 or checker or quadrant or lowcontrast) or random_configurations and (0 or 5) or colour_ingest or device_resident or chunked or stereo_batch or rgbd_frame_chain or rectified_stereo_chain or search_by_projection_map and 3.0 or search_by_bow_kf or two_cameras and (3.0 or True-0) or search_for_triangulation and True or distinctive or bow_transform or undistort_keypoints_bit_exact and 0'
FILES="tests/test_gpu_extractor.py tests/test_gpu_stereo.py tests/test_gpu_frame_ops.py tests/test_gpu_matcher_methods.py tests/test_gpu_matcher_kf.py tests/test_gpu_matcher.py"
for tool in memcheck racecheck; do
    timeout 1500 compute-sanitizer --tool $tool --target-processes all --error-exitcode 66 --print-limit 20 \
        python -m pytest $FILES -m gpu -x -q -k "$SEL" > $OUT/sanitize_$tool.log 2>&1
    echo "$tool rc=$?"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error" $OUT/sanitize_$tool.log | tail -5
done
