"""Per-phase wall-clock split and cProfile of one training step (diagnostic; not a bench line)."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import impflow_b200 as pkg  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'cifar']
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
opt = torch.optim.Adam(params, lr=1e-3, betas=(0.9, 0.99))
n_dims = c * h * w


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


def step(timing=None):
    t0 = sync()
    bucket.zero()
    z, dlogp = model(x, 0)
    logpz = bench.std_normal_logprob(z).reshape(z.size(0), -1).sum(1, keepdim=True)
    bpd = -torch.mean(logpz - dlogp - np.log(256) * n_dims) / n_dims / np.log(2)
    t1 = sync()
    bpd.backward()
    t2 = sync()
    bucket.allreduce_mean()
    torch.nn.utils.clip_grad_norm_(params, 1.)
    opt.step()
    t3 = sync()
    bench.update_lipschitz(pkg, model)
    t4 = sync()
    if timing is not None:
        timing.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))


for _ in range(2):
    step()
tm = []
l0 = pkg._cabi.launch_count()
for _ in range(3):
    step(tm)
print('launches/step', (pkg._cabi.launch_count() - l0) / 3)
tm = np.array(tm) * 1e3
print('ms  forward %.1f  backward %.1f  reduce+clip+adam %.1f  update_lipschitz %.1f' % tuple(tm.mean(0)))
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue()[:9000])
