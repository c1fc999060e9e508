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

from impflow_b200.layers import implicit_block  # noqa: E402
implicit_block.PROBE_MODE['mode'] = os.environ.get('PROBE_MODE', 'device')     # what bench.py times
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


for _ in range(4):
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
# GPU busy time = union of the kernel intervals over all streams; idle = first start .. last end minus busy
def _short(n):
    n = n.replace('void ', '').replace('at::native::', '').replace('(anonymous namespace)::', '')
    return n[:110]


iv = sorted((ev.time_range.start, ev.time_range.end, _short(ev.name)) for ev in prof.events()
            if ev.device_type == torch.autograd.DeviceType.CUDA)
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
gaps = []
pairs = []
late = collections.defaultdict(float)      # idle time by the kernel that ends the gap ("who was late")
late_n = collections.Counter()
prev_name = iv[0][2]
after = collections.defaultdict(float)     # ... and by the kernel that ran before the gap
for s, e, nm in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append(s - cur_e)
        pairs.append((s - cur_e, prev_name, nm))
        late[nm] += s - cur_e
        late_n[nm] += 1
        after[prev_name] += s - cur_e
        cur_s, cur_e = s, e
        prev_name = nm
    else:
        if e > cur_e:
            prev_name = nm
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
span = iv[-1][1] - iv[0][0]
gaps = np.array(gaps) if gaps else np.zeros(1)
print('GPU span %.1f ms, busy (union) %.1f ms, idle %.1f ms in %d gaps (median gap %.1f us, gaps > 20 us: %d totalling %.1f ms)' % (
    span / 1e3, busy / 1e3, (span - busy) / 1e3, len(gaps), float(np.median(gaps)), int((gaps > 20).sum()),
    float(gaps[gaps > 20].sum()) / 1e3))
print('idle time by the kernel that follows the gap:')
for k, v in sorted(late.items(), key=lambda kv: -kv[1])[:16]:
    print('   %8.0f us in %4d gaps (avg %6.1f us) before %s' % (v, late_n[k], v / late_n[k], k))
print('idle time by the kernel that precedes the gap:')
for k, v in sorted(after.items(), key=lambda kv: -kv[1])[:10]:
    print('   %8.0f us after %s' % (v, k))
print('largest gaps (us, kernel before -> kernel after):')
for g, a, b in sorted(pairs, key=lambda t: -t[0])[:45]:
    print('   %7.1f  %s  ->  %s' % (g, a[:70], b[:90]))
