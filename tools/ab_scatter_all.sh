#!/bin/bash
# same-box A/B of scatter library variants over every config-4 mode: tools/ab_scatter_all.sh libA.so libB.so
for rep in 1 2; do
  for lib in "$@"; do
    echo "== $(basename $lib) rep$rep"
    V2V_B200_LIB=$PWD/$lib python tools/time_scatter.py 2>&1 | grep -v ranges
  done
done
