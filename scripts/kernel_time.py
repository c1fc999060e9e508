"""In-situ GPU kernel time per training step by kernel name (torch.profiler / CUPTI); diagnostic."""
import collections
import os
import sys
import time

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

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


def step():
    bucket.zero()
    z, dlogp = model(x, 0)
    logpz = bench.std_normal_logprob(z).reshape(z.size(0), -1).sum(1, keepdim=True)
    bpd = -torch.mean(logpz - dlogp - np.log(256) * n_dims) / n_dims / np.log(2)
    bpd.backward()
    bucket.allreduce_mean()
    torch.nn.utils.clip_grad_norm_(params, 1.)
    opt.step()
    bench.update_lipschitz(pkg, model)


for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
step()
torch.cuda.synchronize()
print('wall ms/step (no profiler): %.1f' % ((time.perf_counter() - t0) * 1e3))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
tot = collections.defaultdict(float)
cnt = collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split('(')[0][:70]
        tot[name] += ev.device_time
        cnt[name] += 1
T = sum(tot.values())
print('GPU kernel time per step: %.1f ms over %d kernels' % (T / 1e3, sum(cnt.values())))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:25]:
    print('%9.0f us %5.1f%% n=%5d avg=%8.1f  %s' % (v, 100 * v / T, cnt[k], v / cnt[k], k))
