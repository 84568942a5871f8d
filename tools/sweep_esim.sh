#!/bin/bash
# Same-box sweep of the throughput kernel's CTA geometries (V2V_ESIM_GEOM, see launch_esim_fast) on 32 config-2 clips.
for noise in philox none; do
  for stats in "--stats" ""; do
    for g in 0 1 2 3; do
      V2V_ESIM_GEOM=$g python tools/profile_esim.py --noise $noise --clips 32 --iters 6 --time $stats
    done
  done
done
