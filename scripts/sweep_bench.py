"""Config 4 timing: compute_ffi + IoU/F1 sweep over N 128x128 patch pairs (resident in HBM)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from rfi_toolbox_b200 import compute_ffi_batch, evaluate_segmentation_batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
g = torch.Generator(device='cuda').manual_seed(7)
truth = torch.rand((n, 128, 128), device='cuda', generator=g) < 0.10
pred = truth ^ (torch.rand((n, 128, 128), device='cuda', generator=g) < 0.02)
data = torch.randn((n, 128, 128), device='cuda', generator=g, dtype=torch.float32).to(torch.complex64)
data = (data + 1j * torch.randn((n, 128, 128), device='cuda', generator=g)) * torch.where(truth, 100.0, 1.0)
for _ in range(2):
    a = compute_ffi_batch(data, pred); b = evaluate_segmentation_batch(pred, truth)
torch.cuda.synchronize(); t0 = time.perf_counter()
a = compute_ffi_batch(data, pred); torch.cuda.synchronize(); t1 = time.perf_counter()
b = evaluate_segmentation_batch(pred, truth); torch.cuda.synchronize(); t2 = time.perf_counter()
px = n * 128 * 128
print(f'{n} pairs: ffi sweep {1e3*(t1-t0):.1f} ms ({px/(t1-t0)/1e9:.1f} Gpix/s), metric sweep {1e3*(t2-t1):.1f} ms ({px/(t2-t1)/1e9:.1f} Gpix/s); mean ffi {a["ffi"].mean():.4f} mean iou {b["iou"].mean():.4f}')
