#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dcn_gpu.py tests/test_edge_cases_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/gpu/dcn_ref_compare.py 15 > gpurun_out/r56_dcn_ref_compare.txt 2>&1; grep "dcn_tc_kernel\|maxdiff vs" gpurun_out/r56_dcn_ref_compare.txt
