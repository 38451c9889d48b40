"""BASELINE config 4 through the C ABI: rfi_pair_sweep over n pairs of 128 x 128 complex64 (kernel only,
device-event time), and evaluate_pairs (kernel + result download + host formulas)."""
import ctypes as C, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from rfi_toolbox_b200 import _native, evaluate_pairs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
lib = _native.load()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
true = torch.rand((n, 128, 128), generator=g, device=dev) < 0.10
pred = true ^ (torch.rand((n, 128, 128), generator=g, device=dev) < 0.02)
data = torch.view_as_complex(torch.randn((n, 128, 128, 2), generator=g, device=dev))
data = (data * (1.0 + 99.0 * true)).contiguous()
res = torch.empty((n, 88), dtype=torch.uint8, device=dev)   # rfi_pair_result_t
st = torch.cuda.current_stream().cuda_stream
def kern():
    _native.check(lib.rfi_pair_sweep(data.data_ptr(), _native.RFI_C64, pred.view(torch.uint8).data_ptr(), true.view(torch.uint8).data_ptr(),
                                     n, 128 * 128, None, res.data_ptr(), st), "pair_sweep")
for _ in range(2):
    kern()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(reps):
    kern()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"rfi_pair_sweep {n} pairs: {ms:.3f} ms  {n * 16384 / ms / 1e6:.1f} Gpix/s  {n * 16384 * 10 / ms / 1e6:.0f} GB/s")
evaluate_pairs(data, pred, true, errors="nan")
t0 = time.perf_counter()
for _ in range(reps):
    r = evaluate_pairs(data, pred, true, errors="nan")
torch.cuda.synchronize()
print(f"evaluate_pairs: {(time.perf_counter() - t0) / reps * 1e3:.3f} ms per sweep; status counts", np.unique(np.frombuffer(res.cpu().numpy().tobytes(), dtype=np.dtype([('f','f8',9),('c','u4',3),('s','i4')]))['s'], return_counts=True))
