"""k_update (csrc/broyden.cu): history-read-once chunked kernel against the two-pass one.
For each shape and chunk setting: n iterations of impflow_broyden_step on fixed random residuals, each timed alone
(CUDA events, L2 flushed); the iterates and the history must be BIT-IDENTICAL across settings.
Usage: python scripts/update_bench.py [B d n_iter] ...   (default: the classifier and bench shapes)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import impflow_b200 as pkg                                   # noqa: E402
from impflow_b200.layers import broyden as _b                # noqa: E402


def run(B, d, T, n_iter, chunk, flush, seed=0):
    cabi = pkg._cabi
    lib = cabi.load()
    lib.impflow_broyden_set_chunk(chunk)
    dev = torch.device('cuda')
    gen = torch.Generator(device='cuda').manual_seed(seed)
    wk = _b._workspace(B, d, T, dev)
    wk.Ut.zero_()
    wk.Vt.zero_()
    gs = [torch.randn(B, d, device=dev, generator=gen) for _ in range(n_iter + 1)]
    wk.xa.copy_(torch.randn(B, d, device=dev, generator=gen))
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    cabi.check(lib.impflow_broyden_begin(vp(wk.xa), vp(gs[0]), vp(wk.xb), vp(wk.low_x), vp(wk.low_g), vp(wk.sample_sq),
                                         vp(wk.low_sq), vp(wk.partial), vp(wk.state), B, d, T, 1e-30, cabi.stream()),
               'begin')
    x_old, xn = wk.xa, wk.xb
    times = []
    for i in range(1, n_iter + 1):
        g_old, gn = gs[i - 1], gs[i]
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cabi.check(lib.impflow_broyden_step(vp(x_old), vp(g_old), vp(xn), vp(gn), vp(wk.Ut), vp(wk.Vt), vp(wk.low_x),
                                            vp(wk.low_g), vp(wk.sample_sq), vp(wk.low_sq), vp(wk.partial),
                                            vp(wk.state), B, d, T, cabi.stream()), 'step')
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3)
        x_old, xn = xn, x_old
    lib.impflow_broyden_set_chunk(-1)
    return times, (xn.clone(), wk.Ut[:, :n_iter].clone(), wk.Vt[:, :n_iter].clone())


def main():
    shapes = [(128, 65536, 15), (128, 16384, 15), (64, 3072, 15)]
    if len(sys.argv) > 3:
        a = [int(t) for t in sys.argv[1:]]
        shapes = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)]
    flush = torch.empty(64 * 1024 * 1024, device='cuda')
    hbm = 6543.1
    for B, d, n in shapes:
        ref = None
        for chunk in (0, -1, 1, 2, 4, 8):
            run(B, d, 30, 2, chunk, flush)          # warm-up
            times, out = run(B, d, 30, n, chunk, flush)
            same = True if ref is None else all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(ref, out))
            if ref is None:
                ref = out
            alg = [(6 + 2 * i) * d * 4.0 * B for i in range(1, n + 1)]
            frac = [a / (t * 1e-6) / 1e9 / hbm for a, t in zip(alg, times)]
            tot = sum(alg) / (sum(times) * 1e-6) / 1e9 / hbm
            print('B=%d d=%d chunk=%2d  bit-identical=%s  us/iter: %s  frac@i=1,4,8,%d: %.2f %.2f %.2f %.2f  overall %.3f'
                  % (B, d, chunk, same, ' '.join('%.0f' % t for t in times), n, frac[0], frac[3], frac[min(7, n - 1)],
                     frac[-1], tot), flush=True)
            del out
        del ref
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
