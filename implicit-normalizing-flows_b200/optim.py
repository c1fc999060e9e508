"""Step tail of the training loop as one fused pass: gradient clipping, the
Adam update of the reference's vendored optimiser (lib/optimizers.py:47-107) and the parameter EMA
(lib/utils.py:140-146) over flat buffers, with the clip norm kept on the device (no host sync).

    bucket = parallel.FlatGradBucket(params)            # gradients live in one flat buffer
    opt = optim.FusedAdam(params, lr=1e-3, betas=(0.9, 0.99), bucket=bucket, max_grad_norm=1.)
    loss.backward(); bucket.allreduce_mean(); opt.step()

`step()` = clip_grad_norm_(params, max_grad_norm) + Adam.step() (+ ema.apply()) of train_img.py:652-658."""
import math

import torch

from . import _cabi, ops
from .parallel import FlatGradBucket

__all__ = ['FusedAdam']


class FusedAdam(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, bucket=None,
                 max_grad_norm=None, ema_decay=None):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        if amsgrad:
            raise NotImplementedError('impflow_b200.optim.FusedAdam: amsgrad is not used by any shipped config')
        # weight_decay is accepted and ignored like in the reference, whose decay line is a no-op
        # (`p.data.add(...)` without the underscore, lib/optimizers.py:105-106)
        params = [p for p in params if p.requires_grad]
        super(FusedAdam, self).__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.bucket = bucket if bucket is not None else FlatGradBucket(params)
        if [id(p) for p in self.bucket.params] != [id(p) for p in params]:
            raise ValueError('FusedAdam: the gradient bucket must hold exactly the optimised parameters, in order')
        self.max_grad_norm = max_grad_norm
        self.ema_decay = ema_decay
        n = self.bucket.flat.numel()
        dev = self.bucket.flat.device
        # parameters re-pointed at one flat buffer (values preserved)
        self.flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, off in zip(params, self.bucket.offsets):        # same (aligned) layout as the gradients
                v = self.flat_p[off:off + p.numel()].view_as(p)
                v.copy_(p.data)
                p.data = v
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        # EMA shadow (lib/utils.py:126-146): the reference initialises it on the FIRST apply(), i.e. from the
        # parameters AFTER the first optimiser step (and after any data-dependent init); later applies blend
        self.ema = torch.zeros_like(self.flat_p) if ema_decay is not None else None
        self._ema_started = False
        self.step_count = 0
        self.last_grad_norm_sq = None        # device scalar of the latest step (sqrt it to log the norm)
        self._ever_grad = [False] * len(self.bucket.params)

    def zero_grad(self, set_to_none=True):
        self.bucket.zero()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self.bucket.gather_strays()
        g = self.bucket.flat
        base = self.flat_p.data_ptr()
        for p, off in zip(self.bucket.params, self.bucket.offsets):
            if p.data_ptr() != base + 4 * off:
                raise RuntimeError('FusedAdam: a parameter no longer lives in the flat buffer (model.to() / .float() / '
                                   'p.data = ... after the optimiser was built); build the optimiser last')
        self.step_count += 1
        group = self.param_groups[0]
        beta1, beta2 = group['betas']
        t = self.step_count
        step_size = group['lr'] * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
        gsq = None
        if self.max_grad_norm is not None:
            gsq = ops._flat_dot(g, g)        # 256 CTAs + a 256-element sum (the bucket length is a multiple of 1024)
            self.last_grad_norm_sq = gsq
        _cabi.check(_cabi.load().impflow_clip_adam_ema(
            _cabi.ptr(self.flat_p), _cabi.ptr(g), _cabi.ptr(self.exp_avg), _cabi.ptr(self.exp_avg_sq),
            _cabi.ptr(self.ema if self._ema_started else None, 'ema', True), g.numel(),
            _cabi.ptr(gsq, 'gnorm_sq', True),
            float(self.max_grad_norm or 0.0), float(step_size), float(beta1), float(beta2), float(group['eps']),
            float(self.ema_decay or 0.0), _cabi.stream()), 'clip_adam_ema')
        if self.ema is not None and not self._ema_started:
            self.ema.copy_(self.flat_p)          # first apply(): copy only (lib/utils.py:140-142)
            self._ema_started = True
        # the kernel wrote through raw pointers: the host caches are keyed on the tensors' versions.  Parameters
        # that received no gradient (the roulette rates geom_p / lamb: the vendored Adam skips them,
        # lib/optimizers.py:70-72) keep value and version — a zero gradient leaves them bit-identical here too —
        # so caches of their host copies (imBlock._rate) stay valid and nothing re-reads them from the device
        # (a parameter that had a gradient in ANY earlier step still moves with its momentum: its version follows)
        self._ever_grad = [e or h for e, h in zip(self._ever_grad, self.bucket.had_grad)]
        torch.autograd.graph.increment_version([p for p, e in zip(self.bucket.params, self._ever_grad) if e])
        return loss

    # ---- checkpointing: the layout of torch.optim / lib.optimizers.Adam state dicts (train_img.py:846,854:
    # 'optimizer_state_dict'), so a checkpoint written by the reference's Adam loads here and vice versa; the EMA
    # shadow travels under the extra key 'impflow_ema' (the reference pickles its EMA object separately)
    def _views(self, flat):
        return [flat[off:off + p.numel()].view_as(p) for p, off in zip(self.bucket.params, self.bucket.offsets)]

    def ema_params(self):
        """EMA shadow of every optimised parameter (views, parameter order), or None before the first step."""
        return self._views(self.ema) if (self.ema is not None and self._ema_started) else None

    def _publish_state(self):
        self.state.clear()
        if self.step_count > 0:
            for p, m, v in zip(self.bucket.params, self._views(self.exp_avg), self._views(self.exp_avg_sq)):
                self.state[p] = {'step': self.step_count, 'exp_avg': m, 'exp_avg_sq': v}

    def state_dict(self):
        self._publish_state()
        sd = super(FusedAdam, self).state_dict()
        ema = self.ema_params()
        sd['impflow_ema'] = None if ema is None else [e.clone() for e in ema]
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        ema = state_dict.pop('impflow_ema', None)
        super(FusedAdam, self).load_state_dict(state_dict)
        steps = []
        for p, m, v in zip(self.bucket.params, self._views(self.exp_avg), self._views(self.exp_avg_sq)):
            st = self.state.get(p)
            if st:
                m.copy_(st['exp_avg'])
                v.copy_(st['exp_avg_sq'])
                steps.append(int(st['step']))
            else:              # the vendored Adam keeps no state for parameters that never had a gradient
                m.zero_()
                v.zero_()
        self.step_count = max(steps) if steps else 0
        self._ever_grad = [bool(self.state.get(p)) for p in self.bucket.params]
        self._publish_state()
        if self.ema is not None:
            self._ema_started = ema is not None
            if ema is not None:
                for dst, src in zip(self._views(self.ema), ema):
                    dst.copy_(src)

    def grad_norm(self):
        """||g|| before clipping of the latest step, as a device scalar (what clip_grad_norm_ returns)."""
        return torch.sqrt(self.last_grad_norm_sq) if self.last_grad_norm_sq is not None else None
