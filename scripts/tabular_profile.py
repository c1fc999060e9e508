"""cProfile of one tabular training step (host-bound path; diagnostic)."""
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
from impflow_b200.layers import implicit_block as ib  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'tabular-power']
ib.PROBE_MODE['mode'] = 'device'
torch.manual_seed(0)
np.random.seed(0)
dev = torch.device('cuda:0')
model = bench.build_mlp_flow(pkg, wl).to(dev)
x = torch.randn(wl['batch'], wl['d'], device=dev)
with torch.no_grad():
    for n_, p_ in model.named_parameters():
        if n_.endswith('weight') and p_.dim() == 2 and p_.requires_grad:
            p_.mul_(30.0 if p_.shape[0] == wl['d'] else 1.0)
    model(x, restore=True)
model.train()
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-3)


def step():
    opt.zero_grad()
    z, dlogp = model(x, torch.zeros(x.shape[0], 1, device=dev))
    loss = -(bench.std_normal_logprob(z).sum(1, keepdim=True) - dlogp).mean()
    loss.backward()
    opt.step()
    bench.update_lipschitz(pkg, model, wl.get('n_lipschitz_iters'))


for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
step()
torch.cuda.synchronize()
print('step wall %.1f ms' % ((time.perf_counter() - t0) * 1e3))
l0 = pkg._cabi.launch_count()
pr = cProfile.Profile()
pr.enable()
step()
torch.cuda.synchronize()
pr.disable()
print('own launches per step', pkg._cabi.launch_count() - l0)
for key in ('tottime', 'cumulative'):
    sio = io.StringIO()
    pstats.Stats(pr, stream=sio).sort_stats(key).print_stats(28)
    print(sio.getvalue()[:5500])

from torch.profiler import ProfilerActivity, profile
import collections
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in rows)
print('GPU kernel time per step: %.1f ms over %d kernels' % (tot / 1e3, sum(e.count for e in rows)))
for e in rows[:18]:
    print('  %8.0f us %5.1f%% n=%5d avg=%6.1f  %s' % (e.self_device_time_total, 100 * e.self_device_time_total / tot, e.count,
                                                    e.self_device_time_total / e.count, e.key[:80]))
