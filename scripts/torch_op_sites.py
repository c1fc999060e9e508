"""Which Python lines of the step issue torch (ATen) CUDA kernels, and how many: the host-bound glue that leaves
the GPU idle between our own kernels.  Diagnostic; python scripts/torch_op_sites.py"""
import collections
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

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


def step():
    bucket.zero()
    z, dlogp = model(x, 0)
    logpz = bench.std_normal_logprob(z).reshape(z.size(0), -1).sum(1, keepdim=True)
    bpd = -torch.mean(logpz - dlogp - np.log(256) * n_dims) / n_dims / np.log(2)
    bpd.backward()
    bucket.allreduce_mean()
    opt.step()
    bench.update_lipschitz(pkg, model)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
sites = collections.Counter()
ops = collections.Counter()
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CPU or not ev.name.startswith('aten::'):
        continue
    if not ev.kernels:          # only ops that launched something themselves
        continue
    n = len(ev.kernels)
    ops[ev.name] += n
    site = 'autograd engine / no python frame'
    for fr in ev.stack or []:
        if ('impflow_b200' in fr or 'implicit-normalizing-flows_b200' in fr or 'bench.py' in fr) and 'scripts/' not in fr:
            site = fr.split('implicit-normalizing-flows_b200/')[-1].split('impflow_b200/')[-1]
            break
    else:
        if ev.stack:
            site = 'other: ' + ' | '.join(f[-60:] for f in ev.stack[:3])
    sites[(site, ev.name)] += n
print('torch-launched kernels per step: %d' % sum(ops.values()))
for k, v in ops.most_common(15):
    print('  %5d  %s' % (v, k))
print('by call site:')
for (site, name), v in sites.most_common(60):
    print('  %5d  %-22s %s' % (v, name, site[:150]))
