"""Layer-1 GEMM shapes of the wider scales (K = 9c padded, N = C): epilogue flavour x CTA-pair x TMA-store sweep
(CUDA events, L2 flushed).  Shows what bounds the short-K launches: `python scripts/gemm_l1_bench.py`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import impflow_b200 as pkg  # noqa: E402

ops = pkg.ops
lib = pkg._cabi.load()
ops.set_gemm_backend('tc')
flush = torch.empty(64 * 1024 * 1024, device='cuda')
names = ['pre', 'act', 'dmul', 'split', 'splitk']
for M, N, K in [(16384, 512, 128), (4096, 512, 448), (65536, 512, 32), (16384, 512, 512)]:
    for flags in [(True, False, False, False, False),     # one fp32 plane, no activation
                  (False, True, False, False, False),     # one fp32 plane, LipSwish
                  (False, False, False, True, False),     # hi/lo planes, LipSwish (forward)
                  (True, False, False, True, False),      # forward + save
                  (False, False, True, True, False),      # vjp: act' product -> planes
                  (True, False, True, False, False)]:     # vjp: act' product -> one fp32 plane
        row = []
        for pair in (1, 0):
            for tma in (1, 0):
                lib.impflow_gemm_tc_set_pair(pair)
                lib.impflow_gemm_tc_set_tma_store(tma)
                t = ops.time_gemm_shape((M, N, K) + flags, reps=5, flush=flush)
                row.append('pair=%d tma=%d %6.1f us' % (pair, tma, t * 1e3))
        lib.impflow_gemm_tc_set_pair(1)
        lib.impflow_gemm_tc_set_tma_store(1)
        tag = '+'.join(n for n, f in zip(names, flags) if f)
        print('M=%6d N=%4d K=%4d %-12s: %s' % (M, N, K, tag, ' | '.join(row)))
