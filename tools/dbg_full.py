import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import v2v_oracle as orc
import v2v_oracle_c as orcc
import v2v_b200 as v2v
from v2v_b200 import _lib
from conftest import synth_video
dev = torch.device('cuda:0')
n, h, w = 121, 480, 640
vid = synth_video("walk", n, h, w, 1234)
pos, neg, std, frac, hstd = 0.21, 0.29, 0.06, 0.0007, 8.0
lut = orc.esim_log_lut()
u0, hot, bn = v2v.philox_fields(n, h, w, base_noise_std=std, hot_pixel_fraction=frac, hot_pixel_std=hstd, seed=5, clip_index_base=17)
u0, hot, bn = u0[0].cpu().numpy(), hot[0].cpu().numpy(), bn[0].cpu().numpy()
def run(nn, flags=0):
    fr = torch.from_numpy(vid[:nn]).to(dev)
    o = v2v.frames_to_voxel(fr, pos, neg, num_bins=1, noise="philox", base_noise_std=std, hot_pixel_fraction=frac, hot_pixel_std=hstd,
                            seed=5, clip_index_base=17, with_stats=True, return_potential=True, kernel_flags=flags)
    return o.voxel[0].reshape(nn - 1, h * w).cpu().numpy(), o.potential[0].cpu().numpy().ravel()
vf, pf = run(n); vg, pg = run(n, _lib.ESIM_FLAG_GENERIC)
ref, pot = orcc.esim_video_to_voxel(vid, pos, neg, 1.0, u0, hot, bn, False, lut, return_state=True)
pot = pot.ravel()
print("pot fast!=generic", int((pf != pg).sum()), "fast!=oracle", int((pf != pot).sum()), "generic!=oracle", int((pg != pot).sum()))
bad = np.flatnonzero(pf != pot)
print("bad pixels", bad[:20].tolist(), "hot", hot.ravel()[bad[:20]].tolist())
print("diffs", (pf[bad[:10]] - pot[bad[:10]]).tolist(), "vals", pf[bad[:10]].tolist(), pot[bad[:10]].tolist())
if len(bad):
    p = int(bad[0])
    for nn in (2, 3, 5, 9, 13, 17, 25, 33, 41, 61, 81, 101, 121):
        _, pk = run(nn)
        r2, p2 = orcc.esim_video_to_voxel(vid[:nn], pos, neg, 1.0, u0, hot, bn[:nn - 1], False, lut, return_state=True)
        print(nn, "pixel", p, pk[p], p2.ravel()[p], "nbad", int((pk != p2.ravel()).sum()))
