"""iResBlock: explicit residual block y = x + g(x) with the same power-series log-det
estimators — API mirror of lib/layers/iresblock.py (block :13-170, estimators :186-270).
In the reference it is only usable by direct call (SequentialFlow passes `restore=`, which this
forward does not accept — SURVEY.md quirk #20); the signature is kept identical."""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from . import implicit_block as ib

__all__ = ['iResBlock']


class iResBlock(nn.Module):

    def __init__(self, nnet, geom_p=0.5, lamb=2., n_power_series=None, exact_trace=False, brute_force=False,
                 n_samples=1, n_exact_terms=2, n_dist='geometric', neumann_grad=True, grad_in_forward=False):
        nn.Module.__init__(self)
        self.nnet = nnet
        self.n_dist = n_dist
        self.geom_p = nn.Parameter(torch.tensor(np.log(geom_p) - np.log(1. - geom_p)))   # float64 Parameter (quirk #12)
        self.lamb = nn.Parameter(torch.tensor(lamb))
        self.n_samples = n_samples
        self.n_power_series = n_power_series
        self.exact_trace = exact_trace
        self.brute_force = brute_force
        self.n_exact_terms = n_exact_terms
        self.grad_in_forward = grad_in_forward
        self.neumann_grad = neumann_grad
        self.register_buffer('last_n_samples', torch.zeros(self.n_samples))
        self.register_buffer('last_firmom', torch.zeros(1))
        self.register_buffer('last_secmom', torch.zeros(1))
        self._inject_n = None
        self._inject_probes = None

    def forward(self, x, logpx=None):
        if logpx is None:
            return x + self.nnet(x)
        g, logdetgrad = self._logdetgrad(x)
        return x + g, logpx - logdetgrad

    def inverse(self, y, logpy=None):
        x = self._inverse_fixed_point(y)
        if logpy is None:
            return x
        return x, logpy + self._logdetgrad(x)[1]

    def _inverse_fixed_point(self, y, atol=1e-5, rtol=1e-5):
        # iresblock.py:69-79
        x, x_prev = y - self.nnet(y), y
        i = 0
        tol = atol + y.abs() * rtol
        while not torch.all((x - x_prev) ** 2 / tol < 1):
            x, x_prev = y - self.nnet(x), x
            i += 1
            if i > 1000:
                break
        return x

    def _logdetgrad(self, x):
        """Returns g(x) and logdet|d(x+g(x))/dx|  (iresblock.py:81-164)."""
        with torch.enable_grad():
            if (self.brute_force or not self.training) and (x.ndimension() == 2 and x.shape[1] == 2):
                x = x.requires_grad_(True)
                g = self.nnet(x)
                jac = ib.batch_jacobian(g, x)
                dets = (jac[:, 0, 0] + 1) * (jac[:, 1, 1] + 1) - jac[:, 0, 1] * jac[:, 1, 0]
                return g, torch.log(torch.abs(dets)).view(-1, 1)

            if self.n_dist == 'geometric':
                p = torch.sigmoid(self.geom_p).item()
                sample_fn = lambda m: ib.geometric_sample(p, m)
                rcdf_fn = lambda k, offset: ib.geometric_1mcdf(p, k, offset)
            else:
                lamb = self.lamb.item()
                sample_fn = lambda m: ib.poisson_sample(lamb, m)
                rcdf_fn = lambda k, offset: ib.poisson_1mcdf(lamb, k, offset)

            n_samples = None
            if self.training and self.n_power_series is not None:
                n_power_series = self.n_power_series
                coeff_fn = lambda k: 1.
            else:
                n_exact = self.n_exact_terms if self.training else 20
                n_samples = np.asarray(self._inject_n) if self._inject_n is not None else sample_fn(self.n_samples)
                n_power_series = int(max(n_samples) + n_exact)
                coeff_fn = lambda k: 1 / rcdf_fn(k, n_exact) * sum(n_samples >= k - n_exact) / len(n_samples)

            if not self.exact_trace:
                vareps = self._inject_probes.to(x) if self._inject_probes is not None else torch.randn_like(x)
                if self.training and self.neumann_grad:
                    estimator_fn = ib.neumann_logdet_estimator
                else:
                    estimator_fn = ib.basic_logdet_estimator
                if self.training and self.grad_in_forward:
                    g, logdetgrad = _MemEffWithOutput.apply(estimator_fn, self.nnet, x, n_power_series, vareps,
                                                            coeff_fn, self.training, *list(self.nnet.parameters()))
                else:
                    x = x.requires_grad_(True)
                    g = self.nnet(x)
                    logdetgrad = estimator_fn(g, x, n_power_series, vareps, coeff_fn, self.training)
            else:
                x = x.requires_grad_(True)
                g = self.nnet(x)
                logdetgrad = ib._exact_trace_series(g, x, n_power_series, coeff_fn)

            if self.training and self.n_power_series is None:
                self.last_n_samples.copy_(ib._upload(torch.as_tensor(np.asarray(n_samples), dtype=torch.float32),
                                                     self.last_n_samples))
                estimator = logdetgrad.detach()
                self.last_firmom.copy_(torch.mean(estimator).to(self.last_firmom))
                self.last_secmom.copy_(torch.mean(estimator ** 2).to(self.last_secmom))
            return g, logdetgrad.view(-1, 1)

    def extra_repr(self):
        return 'dist={}, n_samples={}, n_power_series={}, neumann_grad={}, exact_trace={}, brute_force={}'.format(
            self.n_dist, self.n_samples, self.n_power_series, self.neumann_grad, self.exact_trace, self.brute_force)


class _MemEffWithOutput(torch.autograd.Function):
    """iResBlock flavour of the memory-efficient estimator: also returns g and back-propagates
    grad_g through the kept graph (iresblock.py:186-235)."""

    @staticmethod
    def forward(ctx, estimator_fn, gnet, x, n_power_series, vareps, coeff_fn, training, *g_params):
        ctx.training = training
        with torch.enable_grad():
            x = x.detach().requires_grad_(True)
            g = gnet(x)
            ctx.g = g
            ctx.x = x
            logdetgrad = estimator_fn(g, x, n_power_series, vareps, coeff_fn, training)
            if training:
                grad_x, *grad_params = torch.autograd.grad(logdetgrad.sum(), (x,) + g_params, retain_graph=True,
                                                           allow_unused=True)
                if grad_x is None:
                    grad_x = torch.zeros_like(x)
                ctx.save_for_backward(grad_x, *g_params, *[gp if gp is not None else torch.zeros_like(p)
                                                           for gp, p in zip(grad_params, g_params)])
        return ib.safe_detach(g), ib.safe_detach(logdetgrad)

    @staticmethod
    def backward(ctx, grad_g, grad_logdetgrad):
        if not ctx.training:
            raise ValueError('Provide training=True if using backward.')
        with torch.enable_grad():
            grad_x, *params_and_grad = ctx.saved_tensors
            g, x = ctx.g, ctx.x
            g_params = params_and_grad[:len(params_and_grad) // 2]
            grad_params = params_and_grad[len(params_and_grad) // 2:]
            dg_x, *dg_params = torch.autograd.grad(g, [x] + list(g_params), grad_g, allow_unused=True)
        dL = grad_logdetgrad[0].detach()
        with torch.no_grad():
            grad_x = grad_x * dL
            grad_params = tuple(gp * dL for gp in grad_params)
            if dg_x is not None:
                grad_x = grad_x + dg_x
            grad_params = tuple(dg + gp if dg is not None else gp for dg, gp in zip(dg_params, grad_params))
        return (None, None, grad_x, None, None, None, None) + grad_params
