"""In-situ CUPTI kernel table of one classifier training step (ImplicitResNet18, B = 128); diagnostic."""
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

wl = bench.WORKLOADS['classifier']
dev = torch.device('cuda:0')
torch.manual_seed(0)
np.random.seed(0)
model = bench.build_classifier(pkg, wl).to(dev)
gen = torch.Generator().manual_seed(1)
x, y = bench.synthetic_batch(wl, wl['batch'], gen)
x, y = x.to(dev), y.to(dev)
with torch.no_grad():
    model(x)
model.train()
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-3)


def step():
    opt.zero_grad()
    loss = bench.loss_of(model, wl, x, y)
    loss.backward()
    opt.step()
    bench.update_lipschitz(pkg, model, wl.get('n_lipschitz_iters'))


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
step()
torch.cuda.synchronize()
print('wall ms/step (no profiler): %.1f' % ((time.perf_counter() - t0) * 1e3))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
tot, cnt = collections.defaultdict(float), collections.Counter()
t_min, t_max = 1e30, 0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace('void ', '').replace('at::native::', '')[:100]
        tot[name] += ev.device_time
        cnt[name] += 1
        t_min, t_max = min(t_min, ev.time_range.start), max(t_max, ev.time_range.end)
T = sum(tot.values())
print('GPU kernel time per step: %.1f ms over %d kernels, span %.1f ms' % (T / 1e3, sum(cnt.values()), (t_max - t_min) / 1e3))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:30]:
    print('%9.0f us %5.1f%% n=%5d avg=%8.1f  %s' % (v, 100 * v / T, cnt[k], v / cnt[k], k))
