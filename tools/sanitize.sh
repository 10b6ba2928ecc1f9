#!/bin/bash
# compute-sanitizer over three representative GPU tests.  Writes gpurun_out/sanitizer_<tool>_<test>.txt
mkdir -p gpurun_out
T1="tests/test_gpu_net.py::test_layers_small_net"
T2="tests/test_gpu_tiled_e2e.py::test_infer_tiled_equals_oracle_pipeline_on_gpu_boxes"
T3="tests/test_gpu_nms.py::test_sizes_vs_c_oracle"
for tool in memcheck racecheck synccheck; do
  i=0
  for sel in "$T1" "$T2 -k 200" "$T3 -k 513 or 64 or 5633 or 100"; do
    i=$((i+1))
    out=gpurun_out/sanitizer_${tool}_t${i}.txt
    # shellcheck disable=SC2086
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 7 python -m pytest -x -q -m gpu $sel > $out 2>&1
    echo "$tool t$i rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $out | tr '\n' ' ')"
  done
done
