"""Timeline of the streaming loop on the bench cube (no nsys in the image): CUDA events bracket the statistics kernel,
the writer (current stream) and the counting kernel (side stream) of every step; all offsets are read against ONE
event recorded before the first launch, after the loop has drained.  Shows what DESIGN 5.4 / 6 claim: the counting
kernel of step k runs next to the statistics kernel of a later step, and no gap is left for the host phase.

    python scripts/timeline.py [steps] > profiles/r02_timeline.txt
"""
import sys
from collections import deque
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rfi_toolbox_b200 import Preprocessor  # noqa: E402
from rfi_toolbox_b200.evaluation import metrics as M  # noqa: E402
from rfi_toolbox_b200.utils.synth import device_cube  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    warm = 3
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    cube, _ = device_cube(45, 4, 1024, 1024, seed=1234, device=dev)
    kw = dict(patch_size=128, stretch="SQRT", flag_sigma=5, use_custom_flags=False, augmentation_rotations=4)
    np.random.seed(0)
    ds = Preprocessor(cube, None, magnitude=True).create_dataset(**kw)
    truth = ds.labels ^ (torch.rand(ds.labels.shape, device=dev) < 0.01).to(torch.uint8)
    del ds
    side = M._metrics_stream(dev)

    def submit():
        pre = Preprocessor(cube, None, magnitude=True)
        pre.profile = True
        return pre, pre.create_dataset_async(**kw)

    rows, base = [], None
    pending, counts = deque(), deque()
    total = warm + steps
    for _ in range(min(total, 3)):
        pending.append(submit())
    issued = len(pending)
    for i in range(total):
        if i == warm:
            base = torch.cuda.Event(enable_timing=True)
            base.record()
        pre, pd = pending.popleft()
        np.random.seed(0)
        ds = pd.result()
        cur = torch.cuda.current_stream(dev)
        side.wait_stream(cur)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            e0.record()
            c = M.confusion_counts_async(ds.labels, truth)
            e1.record()
        ds.labels.record_stream(side)
        truth.record_stream(side)
        counts.append(c)
        if issued < total:
            pending.append(submit())
            issued += 1
        if len(counts) > 1:
            counts.popleft().result()
        rows.append((i, pre.events["stats"], pre.events["write"], (e0, e1)))
        del ds, pd
    while counts:
        counts.popleft().result()
    torch.cuda.synchronize()
    print("# bench cube (45 bl x 4 pol x 1024 x 1024 complex64, SQRT, MAD sigma 5, R = 4), streaming loop, 2 calls in flight")
    print("# ms after the event recorded before step 0's host phase; stats / write on the current stream, counts on the side stream")
    print("# step   stats [start, end]      write [start, end]      counts [start, end]     counts overlaps")
    spans = {}
    for i, st, wr, cn in rows:
        if i < warm:
            continue
        t = [base.elapsed_time(e) for e in (*st, *wr, *cn)]
        spans[i - warm] = t
    for k, t in spans.items():
        over = []
        for j, u in spans.items():
            for name, a, b in (("stats", u[0], u[1]), ("write", u[2], u[3])):
                o = min(t[5], b) - max(t[4], a)
                if o > 0.01:
                    over.append(f"{name}[{j}] {o:.2f} ms")
        print(f"  {k:3d}   [{t[0]:7.3f}, {t[1]:7.3f}]   [{t[2]:7.3f}, {t[3]:7.3f}]   [{t[4]:7.3f}, {t[5]:7.3f}]   {', '.join(over)}")
    ks = sorted(spans)[:-3]   # the last three steps drain the queue (nothing is submitted behind them)
    if len(ks) > 2:
        per = (spans[ks[-1]][3] - spans[ks[0]][3]) / (ks[-1] - ks[0])
        busy = np.mean([spans[k][1] - spans[k][0] + spans[k][3] - spans[k][2] for k in ks])
        print(f"# steady state (steps {ks[0]}..{ks[-1]}; the last three steps drain the queue): {per:.3f} ms from writer end to "
              f"writer end; statistics + writer busy {busy:.3f} ms of it")
    # the statistics kernel of a call is enqueued two steps ahead (phase 1 of step k + 2 sits between the writers of k and k + 1)


if __name__ == "__main__":
    main()
