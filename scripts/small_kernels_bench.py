"""Timing of the small layout kernels around the branch GEMMs (diagnostic)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import impflow_b200 as pkg  # noqa: E402

ops = pkg.ops


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (B, H, C, ld) in [(64, 32, 3, 32), (64, 16, 12, 128), (64, 8, 48, 448), (64, 32, 3, 28)]:
    x = torch.randn(B, H, H, C, device='cuda')
    col = torch.randn(B * H * H, 9 * C, device='cuda')
    print('B=%d HW=%d C=%d ld=%d: im2col %.1f us  im2col_split %.1f us  col2im %.1f us' % (
        B, H, C, ld, timeit(lambda: ops.im2col3x3(x, ld=ld)),
        timeit(lambda: ops.im2col3x3_split(x, ld=ld)) if ld % 4 == 0 else float('nan'),
        timeit(lambda: ops.col2im3x3(col, B, H, H, C))))
