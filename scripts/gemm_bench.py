"""Micro-benchmark of the tcgen05 3xTF32 GEMM on the dominant CIFAR shapes (CUDA events, L2 flushed)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import impflow_b200 as pkg
ops = pkg.ops
ops.set_gemm_backend('tc')
flush = torch.empty(64 * 1024 * 1024, device='cuda')
names = ['pre', 'act', 'dmul', 'split', 'splitk']
for key in [(65536, 512, 512, False, True, False, False, False),   # act only
            (65536, 512, 512, False, False, False, True, False),   # planes only
            (65536, 512, 512, True, False, False, True, False),    # forward(save): pre + planes
            (65536, 512, 512, False, False, True, True, False),    # vjp: act' product -> planes
            (65536, 512, 512, False, True, True, True, False),     # tangent: product planes + raw
            (65536, 512, 32, True, False, False, True, False),     # conv1 forward(save)
            (65536, 512, 32, False, False, True, True, False),
            (65536, 32, 512, True, False, False, False, False),    # conv3 (N=27 padded)
            (16384, 512, 512, False, False, False, True, False),
            (16384, 512, 512, False, False, True, True, False),
            (16384, 512, 512, True, False, False, False, False),
            (16384, 512, 128, False, False, False, True, False),
            (16384, 512, 128, False, False, True, True, False),
            (16384, 512, 128, True, False, False, False, False),
            (16384, 128, 512, True, False, False, False, False),
            (4096, 512, 448, False, False, True, True, False),
            (4096, 448, 512, True, False, False, False, False),
            (4096, 512, 512, False, False, False, True, False),
            (512, 512, 65536, True, False, False, False, True),    # conv2 weight gradient (split-K)
            (32, 512, 65536, True, False, False, False, True)]:
    t = ops.time_gemm_shape(key, reps=7, flush=flush)
    M, N, K = key[:3]
    tag = '+'.join(n for n, f in zip(names, key[3:]) if f)
    print('M=%6d N=%4d K=%6d %-22s: %7.1f us  %6.1f TFLOP/s (fp32-equivalent)' % (M, N, K, tag, t * 1e3, 2.0 * M * N * K / t / 1e9))
