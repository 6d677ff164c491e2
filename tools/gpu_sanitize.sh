#!/bin/bash
# compute-sanitizer over the small parity cases of the hand-scheduled kernels (memcheck, then racecheck on shared memory)
TAG=${1:-san}; OUT=gpurun_out; mkdir -p $OUT
SEL="fused_passes or fused_j_dofs or boris_fused_step or strang_resident_sorted or multicell or loop_tail or sorted_fast_path or landau_sobol or reference_sampling_test_on_gpu or two_gaussians or deferred_trailing"
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_2d3v.py tests/test_gpu_sampling.py -m gpu -x -q -k "$SEL" > $OUT/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $OUT/${TAG}_memcheck.log | tail -8
timeout 2400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_2d3v.py -m gpu -x -q -k "fused_passes and 3-2 or boris_fused_step and 32-3-2 or strang_resident_sorted or loop_tail" > $OUT/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $OUT/${TAG}_racecheck.log | tail -8
