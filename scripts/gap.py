"""Where the time between the two phases goes (GPU idle gap) -- run on the GPU box."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from rfi_toolbox_b200 import Preprocessor, evaluate_segmentation, _native
from rfi_toolbox_b200.utils.synth import device_cube
cube, mask = device_cube(45, 4, 1024, 1024, seed=1234, device='cuda')
kw = dict(patch_size=128, stretch='SQRT', flag_sigma=5, use_custom_flags=False)
orig = _native.plan_slots
T = []
def timed(*a, **k):
    t0 = time.perf_counter(); r = orig(*a, **k); T.append(time.perf_counter() - t0); return r
_native.plan_slots = timed
import rfi_toolbox_b200.preprocessing.preprocessor as P
gaps, stats, writes, confs, totals = [], [], [], [], []
for it in range(13):
    np.random.seed(0)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
    e0.record()
    pre = Preprocessor(cube, None, magnitude=True); pre.profile = True
    ds = pre.create_dataset(**kw)
    e1.record()
    m = evaluate_segmentation(ds.labels, ds.labels)
    e2.record(); torch.cuda.synchronize()
    if it >= 3:
        ev = pre.events
        stats.append(ev['stats'][0].elapsed_time(ev['stats'][1])); writes.append(ev['write'][0].elapsed_time(ev['write'][1]))
        gaps.append(ev['stats'][1].elapsed_time(ev['write'][0])); confs.append(e1.elapsed_time(e2)); totals.append(e0.elapsed_time(e2))
    del ds, pre
print('stats %.3f  gap %.3f  write %.3f  eval %.3f  total %.3f ms; plan_slots host %.3f ms' % (
    np.mean(stats), np.mean(gaps), np.mean(writes), np.mean(confs), np.mean(totals), 1e3 * np.mean(T[3:])))
