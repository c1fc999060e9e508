"""Host / device split of BranchProgram.neumann and backward_full on one CIFAR scale-0 branch (diagnostic)."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import impflow_b200 as pkg  # noqa: E402
from impflow_b200.branch_program import compile_branch  # noqa: E402

L = pkg.layers
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 0
c, hw = [(3, 32), (12, 16), (48, 8)][scale]
mk = lambda a, b, k: L.base.get_conv2d(a, b, k, 1, k // 2, coeff=0.9, n_iterations=None, domain=2, codomain=2,
                                       atol=1e-3, rtol=1e-3)
torch.manual_seed(0)
net = torch.nn.Sequential(L.base.Swish(), mk(c, 512, 3), L.base.Swish(), mk(512, 512, 1), L.base.Swish(),
                          mk(512, c, 3)).cuda()
x = torch.randn(64, c, hw, hw, device='cuda')
with torch.no_grad():
    net(x[:2])
prog = compile_branch(net)
w, v = torch.randn_like(x), torch.randn_like(x)


def run(fn, reps=5):
    with torch.no_grad():
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        t_host = (time.perf_counter() - t0) / reps
        torch.cuda.synchronize()
        t_all = (time.perf_counter() - t0) / reps
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        gpu = sum(e.self_device_time_total for e in prof.key_averages())
        n = sum(e.count for e in prof.key_averages())
    return t_host * 1e3, t_all * 1e3, gpu / 1e3, n


with torch.no_grad():
    _, saved = prog.forward_saved(x)
for name, fn in [('forward_saved', lambda: prog.forward_saved(x)),
                 ('forward (no save)', lambda: prog.forward(x)),
                 ('vjp', lambda: prog.vjp(v, saved)),
                 ('backward_full', lambda: prog.backward_full(saved, w)),
                 ('neumann', lambda: prog.neumann(saved, w, v))]:
    th, ta, tg, n = run(fn)
    print('%-18s host-issue %7.2f ms   wall %7.2f ms   GPU busy %7.2f ms over %4d kernels' % (name, th, ta, tg, n))

pr = cProfile.Profile()
with torch.no_grad():
    pr.enable()
    prog.neumann(saved, w, v)
    pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(25)
print(s.getvalue()[:6000])

import collections
for name, fn in [('backward_full', lambda: prog.backward_full(saved, w)), ('neumann', lambda: prog.neumann(saved, w, v))]:
    with torch.no_grad():
        fn()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)
    tot = sum(e.self_device_time_total for e in rows)
    print('---- %s: %.2f ms GPU' % (name, tot / 1e3))
    for e in rows[:16]:
        print('  %8.0f us %5.1f%% n=%3d  %s' % (e.self_device_time_total, 100 * e.self_device_time_total / tot, e.count,
                                                e.key[:70]))


# KERNEL TABLE of one neumann() and one backward_full() call
for name, fn in [('neumann', lambda: prog.neumann(saved, w, v)), ('backward_full', lambda: prog.backward_full(saved, w))]:
    with torch.no_grad():
        fn()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
    print('--- kernels of one %s call (scale %d)' % (name, scale))
    rows = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)
    for e in rows[:14]:
        print('   %8.0f us  n=%3d  %s' % (e.self_device_time_total, e.count, e.key[:70]))
