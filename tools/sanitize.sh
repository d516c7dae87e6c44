#!/bin/bash
# One compute-sanitizer tool per gpurun call (B200_PROFILING.md): smallest cases that cover
# every kernel.  usage: tools/sanitize.sh memcheck|racecheck|synccheck|initcheck
set -e
tool=${1:-memcheck}
compute-sanitizer --tool "$tool" --error-exitcode 9 python -m pytest -q -m gpu -x \
  tests/test_gpu_operators.py tests/test_gpu_edge_cases.py \
  "tests/test_gpu_search_parity.py::test_search_matches_oracle[vo_5x7]" \
  "tests/test_gpu_search_parity.py::test_search_matches_oracle[vo_gps_imu]" \
  "tests/test_gpu_search_parity.py::test_search_matches_oracle[gps_traverse]" \
  "tests/test_gpu_search_parity.py::test_chained_seed_mode"
