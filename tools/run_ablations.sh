#!/bin/bash
# times every ablated library (tools/build_ablations.py) on 32 config-2 clips, Philox + stats, same box
python tools/profile_esim.py --noise philox --clips 32 --iters 6 --time --stats | sed 's/^/full          /'
for n in ${ABLS:-noise trigger cross stats noise_trigger all kpf1 kpf2}; do
  V2V_B200_LIB=v2v_b200/lib/abl_$n.so python tools/profile_esim.py --noise philox --clips 32 --iters 6 --time --stats | sed "s/^/abl_$n /"
done
