#!/bin/bash
# sweep step-kernel launch geometry (graph-chained launches, L2-resident single batch)
for B in 64 128 256; do for E in 1 2 4; do
  echo "block=$B ept=$E: $(B2E_STEP_BLOCK=$B B2E_STEP_EPT=$E python scripts/floor.py 2>&1 | grep -E 'CartPole-v1 (65536|262144)' | tr '\n' ' ')"
done; done
