"""Every synchronising torch call of one training step with its call site (torch.cuda.set_sync_debug_mode).
The solver's own per-iteration state read (C side, inherent to the reference's data-dependent iteration count)
is not a torch call and is not listed.  Diagnostic; python scripts/find_syncs.py [workload]"""
import collections
import os
import sys
import traceback
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import impflow_b200 as pkg  # noqa: E402
from impflow_b200.layers import implicit_block  # noqa: E402

implicit_block.PROBE_MODE['mode'] = 'device'
name = sys.argv[1] if len(sys.argv) > 1 else 'cifar'
wl = bench.WORKLOADS[name]
batch = wl['batch']
dev = torch.device('cuda:0')
torch.manual_seed(0)
np.random.seed(0)
model = bench.build_model(pkg, wl, batch).to(dev)
is_mlp = wl.get('kind') == 'mlp'
if is_mlp:
    x = torch.randn(batch, wl['d'], device=dev)
else:
    c, h, w = wl['input']
    x = torch.rand(batch, c, h, w, device=dev)
with torch.no_grad():
    model(x, restore=True) if not is_mlp else model(x, torch.zeros(batch, 1, device=dev), restore=True)
model.train()
params = [p for p in model.parameters() if p.requires_grad]
bucket = pkg.parallel.FlatGradBucket(params)
opt = pkg.optim.FusedAdam(params, lr=1e-3, betas=(0.9, 0.99), bucket=bucket, max_grad_norm=None if is_mlp else 1.,
                          ema_decay=None if is_mlp else 0.999)


def step():
    bucket.zero()
    if is_mlp:
        z, dlogp = model(x, torch.zeros(batch, 1, device=dev))
        loss = -(bench.std_normal_logprob(z).reshape(batch, -1).sum(1, keepdim=True) - dlogp).mean()
    else:
        z, dlogp = model(x, 0)
        loss = -torch.mean(bench.std_normal_logprob(z).reshape(batch, -1).sum(1, keepdim=True) - dlogp)
    loss.backward()
    bucket.allreduce_mean()
    opt.step()
    bench.update_lipschitz(pkg, model, wl.get('n_lipschitz_iters'))


for _ in range(3):
    step()
torch.cuda.synchronize()
sites = collections.Counter()
orig = warnings.showwarning


def show(message, category, filename, lineno, file=None, line=None):
    if 'synchroniz' in str(message):
        st = [f for f in traceback.extract_stack() if 'implicit-normalizing-flows_b200' in f.filename or f.filename.endswith('bench.py')]
        where = ' <- '.join('%s:%d' % (os.path.basename(f.filename), f.lineno) for f in reversed(st[-3:])) or '%s:%d' % (filename, lineno)
        sites[where] += 1
    else:
        orig(message, category, filename, lineno, file, line)


warnings.showwarning = show
warnings.simplefilter('always')
torch.cuda.set_sync_debug_mode('warn')
step()
torch.cuda.set_sync_debug_mode('default')
torch.cuda.synchronize()
print('synchronising torch calls in one %s step: %d' % (name, sum(sites.values())))
for k, v in sites.most_common(30):
    print('  %4d  %s' % (v, k))
