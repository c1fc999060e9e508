"""Micro-benchmark of the fused branch tile kernel (csrc/branch_fused.cu) at the CIFAR scale-0 shape,
next to the three unfused tcgen05 GEMMs it replaces (CUDA events, L2 flushed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import impflow_b200 as pkg  # noqa: E402

ops = pkg.ops
dev = torch.device('cuda')
M, C, N3 = 65536, 512, 27
flush = torch.empty(64 * 1024 * 1024, device=dev)
x0 = torch.zeros(M, 32, device=dev)
x0[:, :27] = torch.randn(M, 27, device=dev)
W1 = ops.split_tf32(torch.randn(C, 32, device=dev) / 5)
W2 = ops.split_tf32(torch.randn(C, C, device=dev) / 22)
W3 = ops.split_tf32(torch.randn(N3, C, device=dev) / 22)
b1, b2 = torch.randn(C, device=dev), torch.randn(C, device=dev)
beta = torch.full((1,), 0.97, device=dev)
m1, m2 = torch.randn(M, C, device=dev), torch.randn(M, C, device=dev)


def timeit(fn, reps=7):
    ts = []
    for i in range(reps + 1):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


flop = 2.0 * M * (32 * C + C * C + C * 32)
for tag, fn in [
    ('fwd (act)', lambda: ops.branch3_tc(x0, W1, W2, W3, N3, bias1=b1, bias2=b2, act_kind=ops.ACT_LIPSWISH,
                                         beta1=beta, beta2=beta)),
    ('fwd (act, save pre)', lambda: ops.branch3_tc(x0, W1, W2, W3, N3, bias1=b1, bias2=b2,
                                                   act_kind=ops.ACT_LIPSWISH, beta1=beta, beta2=beta,
                                                   save_pre=True)),
    ('vjp (mul)', lambda: ops.branch3_tc(x0, W1, W2, W3, N3, mul1=m1, mul2=m2)),
    ('fwd (no act)', lambda: ops.branch3_tc(x0, W1, W2, W3, N3)),
]:
    t = timeit(fn)
    print('branch3 %-22s: %7.1f us  %6.1f TFLOP/s (fp32-equivalent)' % (tag, t * 1e3, flop / t / 1e9))
