"""Per-phase wall-clock split of one training step with a device sync at every phase boundary
(diagnostic; not a bench line).  Phases are nested functions of the hot path, patched with timers."""
import collections
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import impflow_b200 as pkg  # noqa: E402
from impflow_b200.layers import implicit_block as ib  # noqa: E402
from impflow_b200 import branch_program as bp  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'cifar']
batch = wl['batch']
dev = torch.device('cuda:0')
torch.manual_seed(0)
np.random.seed(0)
ib.PROBE_MODE['mode'] = 'device'
ib.OVERLAP['on'] = os.environ.get('OVERLAP', '0') == '1'      # one stream: the per-phase timers sync the device
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

ACC = collections.OrderedDict()
CNT = collections.Counter()
LCH = collections.Counter()
ON = {'on': False}
STACK = []


def timed(name, fn):
    def wrapper(*a, **k):
        if not ON['on']:
            return fn(*a, **k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        l0 = pkg._cabi.launch_count()
        STACK.append(0.0)
        out = fn(*a, **k)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        inner = STACK.pop()
        if STACK:
            STACK[-1] += dt
        ACC[name] = ACC.get(name, 0.0) + dt - inner          # exclusive time
        CNT[name] += 1
        LCH[name] += pkg._cabi.launch_count() - l0
        return out
    return wrapper


ib.RootFind.broyden_find_root = staticmethod(timed('fwd solve (RootFind)', ib.RootFind.broyden_find_root))
ib.branch_apply = timed('re-attach branch_apply', ib.branch_apply)
ib.MemoryEfficientLogDetEstimator.payload = staticmethod(
    timed('logdet estimator fwd (chain + neumann)', ib.MemoryEfficientLogDetEstimator.payload))
ib.imBlock.Backward.backward = staticmethod(timed('implicit backward solve', ib.imBlock.Backward.backward))
ib._BranchApply.backward = staticmethod(timed('re-attach backward_full', ib._BranchApply.backward))
bp.BranchProgram.neumann = timed('  of which neumann sweeps', bp.BranchProgram.neumann)
bp.BranchProgram._prep = timed('  weight prep (_prep)', bp.BranchProgram._prep)
if os.environ.get('FINE'):
    bp.BranchProgram._wgrad_gemm_layout = timed('    wgrad (transposes + split + split-K GEMM)', bp.BranchProgram._wgrad_gemm_layout)
    bp.BranchProgram._apply = timed('    _apply (GEMM / conv layer launches)', bp.BranchProgram._apply)
    bp.BranchProgram._finish_param_grads = timed('    _finish_param_grads', bp.BranchProgram._finish_param_grads)
    bp.ops.act_second = timed('    act_second', bp.ops.act_second)
    bp.ops.act_beta_grad = timed('    act_beta_grad', bp.ops.act_beta_grad)
    bp.ops.colsum = timed('    colsum', bp.ops.colsum)
    bp.BranchProgram._ains = timed('    _ains', bp.BranchProgram._ains)
ib.imBlock.forward = timed('imBlock.forward other', ib.imBlock.forward)
bench.update_lipschitz = timed('update_lipschitz', bench.update_lipschitz)
opt.step = timed('adam', opt.step)


def step():
    bucket.zero()
    z, dlogp = model(x, 0)
    logpz = bench.std_normal_logprob(z).reshape(z.size(0), -1).sum(1, keepdim=True)
    bpd = -torch.mean(logpz - dlogp - np.log(256) * n_dims) / n_dims / np.log(2)
    bpd.backward()
    bucket.allreduce_mean()
    opt.step()
    bench.update_lipschitz(pkg, model)


step_t = timed('step other (loss, actnorm, squeeze, clip, autograd glue)', step)
for _ in range(2):
    step()
ON['on'] = True
ms0 = torch.cuda.memory_stats()
N = 3
for _ in range(N):
    step_t()
ms1 = torch.cuda.memory_stats()
for k in ('num_alloc_retries', 'num_device_alloc', 'num_device_free', 'allocation.all.allocated', 'segment.all.allocated'):
    print('allocator %-28s +%d' % (k, ms1.get(k, 0) - ms0.get(k, 0)))
print('reserved GB %.1f  peak allocated GB %.1f' % (ms1['reserved_bytes.all.current'] / 2**30, ms1['allocated_bytes.all.peak'] / 2**30))
tot = sum(ACC.values())
print('synced wall per step: %.1f ms' % (tot / N * 1e3))
for k, v in sorted(ACC.items(), key=lambda kv: -kv[1]):
    print('  %-58s %7.2f ms  %5.1f%%  calls/step %5.1f  own launches/step %6.1f' %
          (k, v / N * 1e3, 100 * v / tot, CNT[k] / N, LCH[k] / N))

if os.environ.get('CPROF'):
    import cProfile
    import io
    import pstats
    ON['on'] = False
    step()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    step()
    pr.disable()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print('host issue time of one step under cProfile: %.1f ms' % (t_host * 1e3))
    for key in ('tottime', 'cumulative'):
        sio = io.StringIO()
        pstats.Stats(pr, stream=sio).sort_stats(key).print_stats(40)
        print(sio.getvalue()[:7000])
