"""One tcgen05 GEMM shape a few times (for ncu captures): python scripts/gemm_one.py M N K [dmul] [split]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import impflow_b200 as p  # noqa: E402

ops = p.ops
ops.set_gemm_backend('tc')
M, N, K = [int(v) for v in sys.argv[1:4]]
flags = sys.argv[4:]
A = torch.randn(M, K, device='cuda')
B = torch.randn(N, K, device='cuda') / K ** 0.5
As, Bs = ops.split_tf32(A), ops.split_tf32(B)
dm = torch.randn(M, N, device='cuda') if 'dmul' in flags else None
for _ in range(4):
    ops.gemm_nt(A, B, None, act_kind=ops.ACT_MULTIPLIER if dm is not None else ops.ACT_NONE, want_pre='split' not in flags,
                dmul_pre=dm, A_split=As, B_split=Bs, want_split='split' in flags)
torch.cuda.synchronize()
print('ok')
