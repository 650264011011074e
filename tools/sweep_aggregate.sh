#!/bin/bash
# the MOLCLR_* tuning switches exist only in the debug-switch build of the library
python -m molclr_b200.build --debug-switches > /dev/null && export MOLCLR_B200_LIB=$PWD/molclr_b200/libmolclr_b200_dbg.so
# A/B sweep of the tile-kernel knobs (ring depth, rows per slot, streaming stores, blocked tiles)
for cfg in "3 3 0 0" "2 3 0 0" "3 2 0 0" "2 4 0 0" "3 4 0 0" "2 5 0 0" "2 6 0 0" "3 3 0 1" "4 2 0 1" "2 4 0 1" "3 3 1 1" "3 3 1 0"; do
  set -- $cfg
  echo "== stages=$1 rows=$2 store_cs=$3 blocked=$4"
  MOLCLR_AGG_STAGES=$1 MOLCLR_AGG_ROWS=$2 MOLCLR_AGG_STORE_CS=$3 MOLCLR_AGG_BLOCKED=$4 ONLY_AGG=1 timeout 120 python tools/bench_rowwise.py 2>&1 | grep "tile kernel"
done
