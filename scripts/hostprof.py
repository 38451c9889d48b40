import sys, time, cProfile, pstats, numpy as np, torch
sys.path.insert(0,'/root/repo')
from rfi_toolbox_b200 import Preprocessor, evaluate_segmentation
from rfi_toolbox_b200.utils.synth import device_cube
cube,mask=device_cube(45,4,1024,1024,seed=1,device='cuda')
kw=dict(patch_size=128,stretch='SQRT',flag_sigma=5,use_custom_flags=False)
def step():
    np.random.seed(0)
    pre=Preprocessor(cube,None,magnitude=True); ds=pre.create_dataset(**kw)
    m=evaluate_segmentation(ds.labels, ds.labels)
    return ds
for _ in range(3): step()
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(10): step()
torch.cuda.synchronize(); print('ms/step', (time.perf_counter()-t0)*100)
pr=cProfile.Profile(); pr.enable()
for _ in range(10): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
