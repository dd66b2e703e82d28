#!/bin/bash
mkdir -p gpurun_out
{
echo "== no L2 hint on host gathers"; python tools/e2e_zero_copy_probe.py
echo "== 64-byte L2 hint kept on host gathers"; SMLVM_GATHER_HOST_HINT=1 python tools/e2e_zero_copy_probe.py
} 2>&1 | grep -E "==|chunks (16|32)" | tee gpurun_out/e2e_zero_copy2.log
