"""Fixed cost and per-iteration cost of the power-iteration kernels (CUDA events; diagnostic)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import impflow_b200 as pkg  # noqa: E402

ops = pkg.ops


def timed(fn, reps=9):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


torch.manual_seed(0)
for (co, ci, h, w) in [(512, 3, 32, 32), (3, 512, 32, 32), (512, 12, 16, 16), (48, 512, 8, 8)]:
    W = torch.randn(co, ci, 3, 3, device='cuda') / (9 * ci) ** 0.5
    u = torch.nn.functional.normalize(torch.randn(co * h * w, device='cuda'), dim=0)
    v = torch.nn.functional.normalize(torch.randn(ci * h * w, device='cuda'), dim=0)
    row = []
    for n_it in (0, 1, 2, 5):
        for want_D in (False, True):
            row.append('%d it%s %6.1f us' % (n_it, '+D' if want_D else '  ',
                                             timed(lambda: ops.sn_power_iter_conv(W, u, v, h, w, n_it, 0.0, 0.0, want_D=want_D))))
    print('conv3x3 %3d->%3d %2dx%2d: ' % (ci, co, h, w) + ' | '.join(row))
W = torch.randn(512, 512, device='cuda') / 512 ** 0.5
u = torch.nn.functional.normalize(torch.randn(512, device='cuda'), dim=0)
v = torch.nn.functional.normalize(torch.randn(512, device='cuda'), dim=0)
print('dense 512x512: ' + ' | '.join('%d it %6.1f us' % (n, timed(lambda: ops.sn_power_iter(W, u, v, n, 0.0, 0.0))) for n in (1, 2, 5)))
