#!/bin/bash
mkdir -p gpurun_out
./tools/_dev/write_probe 2>&1 | tee gpurun_out/write_probe.log
python tools/topk_stagger_sweep.py 2>&1 | grep STAGGER | tee gpurun_out/topk_after.log
python -m pytest tests/test_topk_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/topk_tests.log
