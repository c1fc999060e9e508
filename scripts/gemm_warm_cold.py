"""Cold (L2 flushed) against warm (operands just written, as inside a step) timings of the mid-size tcgen05 GEMM shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import impflow_b200 as pkg  # noqa: E402

ops = pkg.ops
ops.set_gemm_backend('tc')
flush = torch.empty(64 * 1024 * 1024, device='cuda')
small = torch.empty(1024, device='cuda')
shapes = [(16384, 512, 512, False, False, False, True, False), (16384, 108, 512, True, False, False, False, False),
          (4096, 432, 512, True, False, False, False, False), (16384, 512, 128, False, False, True, True, False),
          (16384, 512, 128, True, False, False, True, False), (4096, 512, 448, False, False, True, True, False),
          (65536, 512, 512, False, False, True, True, False)]
for key in shapes:
    cold = ops.time_gemm_shape(key, reps=7, flush=flush)
    warm = ops.time_gemm_shape(key, reps=7, flush=small)
    M, N, K = key[:3]
    print('M=%6d N=%4d K=%4d flags=%s: cold %6.1f us  warm %6.1f us  (%5.1f / %5.1f TFLOP/s)' % (
        M, N, K, ''.join('1' if f else '0' for f in key[3:]), cold * 1e3, warm * 1e3, 2.0 * M * N * K / cold / 1e9,
        2.0 * M * N * K / warm / 1e9))
