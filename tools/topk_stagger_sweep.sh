#!/bin/bash
# sweep of the k-best output stagger (groups, percent of the computed slot); run on the GPU box
mkdir -p gpurun_out
for s in 0 2,100 3,100 4,50 4,100 4,150 6,100 8,100 8,60 19,100 19,60; do
  SMLVM_TOPK_STAGGER=$s python tools/topk_stagger_sweep.py 2>&1 | grep STAGGER
done | tee gpurun_out/topk_stagger.log
