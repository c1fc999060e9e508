"""cProfile of the HOST side of one CIFAR-shape training step, main thread and autograd thread separately
(the autograd engine's worker thread is created by C++: its profiler is switched on from inside the first backward
function that runs on it).  Diagnostic for the host-bound stretches of the step."""
import cProfile
import io
import os
import pstats
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import impflow_b200 as pkg  # noqa: E402
from impflow_b200.layers import implicit_block  # noqa: E402

implicit_block.PROBE_MODE['mode'] = 'device'
wl = bench.WORKLOADS['cifar']
batch = wl['batch']
dev = torch.device('cuda:0')
torch.manual_seed(0)
np.random.seed(0)
model = bench.build_model(pkg, wl, batch).to(dev)
c, h, w = wl['input']
x = torch.rand(batch, c, h, w, device=dev)
with torch.no_grad():
    model(x, restore=True)
model.train()
params = [p for p in model.parameters() if p.requires_grad]
bucket = pkg.parallel.FlatGradBucket(params)
opt = pkg.optim.FusedAdam(params, lr=1e-3, betas=(0.9, 0.99), bucket=bucket, max_grad_norm=1., ema_decay=0.999)
n_dims = c * h * w

PROF = {'on': False, 'bwd': None}
tls = threading.local()


def hook_backward(cls):
    orig = cls.backward

    def wrapped(ctx, *a):
        if PROF['on'] and not getattr(tls, 'enabled', False) and threading.current_thread() is not threading.main_thread():
            PROF['bwd'] = cProfile.Profile()
            PROF['bwd'].enable()
            tls.enabled = True
        return orig(ctx, *a)
    cls.backward = staticmethod(wrapped)


for cls in (implicit_block._BranchApply, implicit_block.imBlock.Backward, implicit_block.MemoryEfficientLogDetEstimator):
    hook_backward(cls)


def step():
    bucket.zero()
    z, dlogp = model(x, 0)
    logpz = bench.std_normal_logprob(z).reshape(z.size(0), -1).sum(1, keepdim=True)
    bpd = -torch.mean(logpz - dlogp - np.log(256) * n_dims) / n_dims / np.log(2)
    t1 = time.perf_counter()
    bpd.backward()
    t2 = time.perf_counter()
    bucket.allreduce_mean()
    opt.step()
    bench.update_lipschitz(pkg, model)
    return t1, t2


for _ in range(6):
    step()
torch.cuda.synchronize()
times = []
for _ in range(5):
    t0 = time.perf_counter()
    t1, t2 = step()
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    times.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0))
print('host ms per step: forward issue %.1f, backward issue %.1f, tail issue %.1f, wait for the GPU %.1f, total %.1f'
      % tuple(1e3 * np.mean([t[i] for t in times]) for i in range(5)))
main = cProfile.Profile()       # Python 3.12: one profiler at a time (sys.monitoring): the two threads in two steps
main.enable()
step()
main.disable()
torch.cuda.synchronize()
PROF['on'] = True
step()
torch.cuda.synchronize()
PROF['on'] = False
bwd_thread_profile = PROF['bwd']


# the profiler of the autograd thread has to be switched off from that thread: one more backward does it
class _Off(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.clone()

    @staticmethod
    def backward(ctx, g):
        if PROF['bwd'] is not None:
            PROF['bwd'].disable()
        return g


t_ = torch.ones(1, device=dev, requires_grad=True)
_Off.apply(t_).sum().backward()
for name, pr in (('MAIN THREAD', main), ('AUTOGRAD THREAD', PROF['bwd'])):
    if pr is None:
        continue
    for key in ('tottime', 'cumulative'):
        sio = io.StringIO()
        pstats.Stats(pr, stream=sio).sort_stats(key).print_stats(32)
        print('=====', name, key)
        print(sio.getvalue()[:7000])
