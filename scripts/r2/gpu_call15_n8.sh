#!/bin/bash
# round 2, 8 GPUs: row-block pipeline on the products shape (blocks 1 / 2 / 4 / 8, FP32 and BF16 operand), adaptive item size
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/r2/inplace_sweep.py --refs 0 --row-blocks 1 2 4 8 --steps 20 2> gpurun_out/r2_c15_sweep.err | grep '^{' > gpurun_out/r2_rowblock_sweep_8.jsonl; echo "sweep rc=$?"; cat gpurun_out/r2_rowblock_sweep_8.jsonl; tail -3 gpurun_out/r2_c15_sweep.err
