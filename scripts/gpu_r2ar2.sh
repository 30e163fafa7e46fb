#!/bin/bash
# round 2, pass ar: which ResnetBlocks fork their res_conv (rows = B*H*W of the block's map), three repeats at 64 samples
cd "$(dirname "$0")/.."
for rep in 1 2 3; do
  echo "B=64 in line:"; DD_NO_FORK=1 python scripts/step_n.py 64 100 2>&1 | tail -1
  echo "B=64 fork<=32768 rows:";   DD_FORK_MAX_ROWS=32768 python scripts/step_n.py 64 100 2>&1 | tail -1
  echo "B=64 fork all:"; DD_FORK_MAX_ROWS=1000000 python scripts/step_n.py 64 100 2>&1 | tail -1
done
