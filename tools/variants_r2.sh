#!/bin/bash
# scratch: single-member timeline experiments
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" STAMP_TAG=$tag timeout 300 python tools/stamps_report.py > gpurun_out/stamps_$tag.json 2> gpurun_out/stamps_$tag.err; cat gpurun_out/stamps_$tag.json; }
run base WGK_X=0
run reuse8 WGK_REUSE_EVERY=8
run prio WGK_NODE_PRIORITY=-1
run fused0 WGK_LEVEL_TASKS=fused0
run fused0_reuse8 WGK_LEVEL_TASKS=fused0 WGK_REUSE_EVERY=8
run all3 WGK_LEVEL_TASKS=fused0 WGK_REUSE_EVERY=8 WGK_NODE_PRIORITY=-1
timeout 600 python -m pytest tests/test_host_library.py -x -q -m gpu > gpurun_out/r2h_pytest.log 2>&1
tail -3 gpurun_out/r2h_pytest.log
