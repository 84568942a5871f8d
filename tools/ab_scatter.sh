#!/bin/bash
# same-box A/B of scatter library variants: tools/ab_scatter.sh v2v_b200/lib/ab_*.so
for rep in 1 2; do
  for lib in "$@"; do
    echo -n "$(basename $lib) rep$rep: "
    V2V_B200_LIB=$PWD/$lib python tools/time_scatter.py 2>&1 | grep "bins=5 h5_interp default"
  done
done
