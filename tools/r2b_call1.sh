#!/bin/bash
mkdir -p gpurun_out
{
SMLVM_BWD_SLOTS_FULL_CLEAR=1 TAG=fullclear python tools/step_time.py
TAG=sparsereset python tools/step_time.py
SMLVM_BWD_SLOTS_FULL_CLEAR=1 TAG=fullclear python tools/step_time.py
TAG=sparsereset python tools/step_time.py
} 2>&1 | tee gpurun_out/step_ab.log
python -m pytest tests/test_hotpath_gpu.py tests/test_fuzz_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/hotpath_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:topk_blk -c 2 -o gpurun_out/topk_cfg python tools/prof_topk_cfg.py > gpurun_out/ncu_topk_cfg.log 2>&1
