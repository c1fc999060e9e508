"""Micro-benchmark of the tcgen05 3xTF32 GEMM on the dominant CIFAR shapes (CUDA events, L2 flushed)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import impflow_b200 as pkg
ops = pkg.ops
lib = pkg._cabi.load()
ops.set_gemm_backend('tc')
flush = torch.empty(64 * 1024 * 1024, device='cuda')
def run(M, N, K, wide, want_split=True, reps=10):
    lib.impflow_gemm_tc_set_wide_tiles(wide)
    A = torch.randn(M, K, device='cuda'); B = torch.randn(N, K, device='cuda') / K ** 0.5
    As, Bs = ops.split_tf32(A), ops.split_tf32(B)
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm_nt(A, B, None, act_kind=ops.ACT_RELU, want_pre=False, want_act=not want_split, A_split=As, B_split=Bs, want_split=want_split)
        e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print('M=%6d N=%4d K=%4d wide=%d split_out=%d : %7.1f us  %6.1f TFLOP/s (fp32-equivalent)' % (M, N, K, wide, want_split, ms * 1e3, 2.0 * M * N * K / ms / 1e9))
for wide in (0, 1):
    run(65536, 512, 512, wide)
    run(65536, 512, 512, wide, want_split=False)
    run(16384, 512, 512, wide)
    run(4096, 512, 512, wide)
    run(65536, 512, 32, wide)
run(65536, 27, 512, 1, want_split=False)
