#!/bin/bash
# geometry sweep of the cluster-resident walk (C4): runs per GPU x (NW, TW, CS)
for R in 64 8; do
  echo "== default plan R=$R"; BINEST_PLAN_DEBUG=1 python scripts/walk_bench.py C4 $R 4 2>&1 | grep -v "^resident plan:" | head -3
  for NW in 16 8; do for TW in 1 2 4; do for CS in 2 4 8 16; do
    echo "-- R=$R NW=$NW TW=$TW CS=$CS"
    BINEST_RES_NW=$NW BINEST_RES_TW=$TW BINEST_RES_CS=$CS timeout 60 python scripts/walk_bench.py C4 $R 4 2>&1 | head -1 | cut -c1-200
  done; done; done
done
