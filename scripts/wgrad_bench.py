"""CUDA-event timing of the MN-major weight-gradient kernel at the CIFAR shapes (L2 flushed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import impflow_b200 as pkg  # noqa: E402

flush = torch.empty(64 * 1024 * 1024, device='cuda')
for M, N1, N2 in [(65536, 512, 512), (16384, 512, 512), (4096, 512, 512), (65536, 512, 32), (16384, 512, 128)]:
    t = pkg.ops.time_wgrad_shape(('wgrad', M, N1, N2), reps=7, flush=flush)
    print('wgrad pixels=%6d N1=%4d N2=%4d: %7.1f us  %6.1f TFLOP/s' % (M, N1, N2, t * 1e3, 2.0 * M * N1 * N2 / t / 1e9))
