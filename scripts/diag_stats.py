import sys, numpy as np, torch
sys.path.insert(0, '.')
from rfi_toolbox_b200 import Preprocessor
from rfi_toolbox_b200.utils.synth import device_cube
nbl = int(sys.argv[1]) if len(sys.argv) > 1 else 45
cube, mask = device_cube(nbl, 4, 1024, 1024, seed=1234, device='cuda')
kw = dict(patch_size=128, stretch='SQRT', flag_sigma=5, use_custom_flags=False)
np.random.seed(0)
pre = Preprocessor(cube, None, magnitude=True); pre.profile = True
for _ in range(3):
    ds = pre.create_dataset(**kw)
torch.cuda.synchronize()
st = pre.last_tile_stats.cpu().numpy().view(np.int32).reshape(-1, 22)
route = st[:, 17]
print('tiles', len(route), 'route counts', {int(k): int(((route & 255) == k).sum()) for k in np.unique(route & 255)}, 'fallback reasons', {int(k): int(((route >> 8) == k).sum()) for k in np.unique(route >> 8)})
print('stats ms', pre.events['stats'][0].elapsed_time(pre.events['stats'][1]), 'write ms', pre.events['write'][0].elapsed_time(pre.events['write'][1]))
