import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
from rfi_toolbox_b200 import Preprocessor
from rfi_toolbox_b200.utils.synth import device_cube
def run(nbl, C, T, P, stretch, **kw):
    cube, mask = device_cube(nbl, 4, C, T, seed=1, device='cuda')
    args = dict(patch_size=P, stretch=stretch, flag_sigma=5, use_custom_flags=False); args.update(kw)
    for _ in range(2):
        np.random.seed(0); pre = Preprocessor(cube, None, magnitude=True); pre.profile=True; ds = pre.create_dataset(**args)
    torch.cuda.synchronize()
    t0=time.perf_counter(); np.random.seed(0); pre = Preprocessor(cube, None, magnitude=True); pre.profile=True; ds = pre.create_dataset(**args); torch.cuda.synchronize(); dt=time.perf_counter()-t0
    st = pre.events['stats'][0].elapsed_time(pre.events['stats'][1]); wr = pre.events['write'][0].elapsed_time(pre.events['write'][1])
    print(f'{nbl}bl {C}x{T} P={P} {stretch}: total {dt*1e3:.2f} ms  stats {st:.2f} write {wr:.2f}  -> {cube.numel()/dt/1e9:.2f} Gpix/s, kept {len(ds)}')
    del ds, pre, cube
run(8,1024,1024,128,'SQRT')
run(8,1024,1024,256,'SQRT')
run(8,1024,1024,512,'SQRT')
run(4,1024,4096,256,'SQRT', flag_sigma=3)
run(4,4096,2048,128,'LOG10')
run(8,1000,1000,128,'SQRT')
