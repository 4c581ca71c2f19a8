import numpy as np, sys
sys.path.insert(0, '.')
import gonzag_b200 as gzb
from oracle import gz_oracle as orc
g = dict(np.load('tests/golden/golden_regional.npz'))
lat, lon = g['lat'].astype(np.float64), g['lon'].astype(np.float64)
rd, r = float(g['rd']), int(g['r'])
bt = gzb.BilinTrack(g['ysat'], g['xsat'], lat, lon, k_ew_per=-1, rd_found_km=rd, np_box_r=r)
ctx = bt.ctx; ctx.set_mask(g['mask'])
v_np, v_bl = ctx.interp(g['tsat'], bt.SM, bt.WB, g['tmod'], g['f32'])
bad = np.where(~np.isclose(v_bl, g['ssh_bl_data'], rtol=1e-6, atol=1e-9))[0]
for k in bad:
    print('k', k, 'gpu', v_np[k], v_bl[k], 'ref', g['ssh_np_data'][k], g['ssh_bl_data'][k])
    print(' NP', bt.NP[k], 'SM', bt.SM[k].tolist(), 'WB gpu', bt.WB[k], 'WB ref', g['WB'][k], 'sum', bt.WB[k].sum(), g['WB'][k].sum(), np.sum(list(g['WB'][k])))
    print(' t', g['tsat'][k], g['tmod'])
    print(' maskvals', [g['mask'][j,i] for j,i in bt.SM[k]])
k = 2067
for label, msk in (('orig', g['mask']), ('ones', np.ones_like(g['mask']))):
    ctx.set_mask(msk)
    a, b = ctx.interp(g['tsat'][k:k+1], bt.SM[k:k+1], bt.WB[k:k+1], g['tmod'], g['f32'])
    print(label, a, b)
    sm2 = bt.SM[k:k+1].copy(); sm2[0,2] = [1,22]; sm2[0,3] = [1,23]
    a, b = ctx.interp(g['tsat'][k:k+1], sm2, bt.WB[k:k+1], g['tmod'], g['f32'])
    print(label, 'row1 only', a, b)
print('mask row0[:30]', g['mask'][0,:30], 'sum row 0', g['mask'][0].sum())
