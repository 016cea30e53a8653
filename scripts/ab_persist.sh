#!/bin/bash
# A/B of the trunk kernel's tile plans on one B200: AZB_NNG_PERSIST=1 (one wave of persistent CTAs) vs 0 (whole waves)
mkdir -p gpurun_out
for rep in 1 2; do
for p in ${AB_MODES:-0 1}; do
  echo "AZB_NNG_PERSIST=$p"
  for cfg in "connect4 bf16x2 6960" "connect4 bf16x2 8192" "connect4 bf16x2 2048" "connect4 bf16x2 600" "brandubh bf16x2 3915" "brandubh bf16x2 4096" "connect4 fp16 6960" "connect4_64 bf16x2 6960"; do
    AZB_NNG_PERSIST=$p timeout 120 python scripts/nn_once.py $cfg 50 2>&1 | tail -1
  done
done
done
