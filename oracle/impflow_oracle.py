"""CPU oracle for the ImpFlow hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file is a CPU (torch-on-host) restatement of the reference algorithm for the
path named in BASELINE.json: the batched Broyden root solve, the Russian-roulette
power-series log-det estimators, the induced-2-norm weight rescale and the two
Lipschitz activations.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package (``implicit-normalizing-flows_b200``) never does.

Parity status: the reference ships NO golden vectors or tests, so
this oracle is pinned against outputs of the reference itself, generated in the
build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference/lib`` unmodified behind two import shims) and committed under
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every function
below against those fixtures.

The arithmetic primitives (GEMM, conv, vjp) are PyTorch ATen on the CPU exactly
as in the reference (README.md:28 pins "PyTorch 1.4"; here torch 2.11) — the
reference has no arithmetic of its own below that level.

Every function cites the reference lines (relative to /root/reference) it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# activations  (lib/layers/base/activations.py:7-12, 64-71)
# --------------------------------------------------------------------------------------


def sin_act(x: torch.Tensor) -> torch.Tensor:
    """``Sin.forward``: sin(2*pi*x) / pi * 0.5  (activations.py:11-12)."""
    return torch.sin(2.0 * math.pi * x) / math.pi * 0.5


def lipswish(x: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    """``Swish.forward``: x * sigmoid(x * softplus(beta)) / 1.1  (activations.py:70-71)."""
    return (x * torch.sigmoid(x * F.softplus(beta))) / 1.1


# --------------------------------------------------------------------------------------
# induced 2-norm rescale  (lib/layers/base/mixed_lipschitz.py:85-132, 276-386, 414-444)
# --------------------------------------------------------------------------------------


def _l2n(t: torch.Tensor) -> torch.Tensor:
    """``normalize_u/_v`` for (co)domain 2 = F.normalize(p=2, dim=0) (mixed_lipschitz.py:414-444)."""
    return F.normalize(t, p=2, dim=0)


def _tol_reached(u, old_u, v, old_v, atol, rtol) -> bool:
    """Early-exit test incl. the signed ``torch.max(u)`` quirk (mixed_lipschitz.py:114-120)."""
    err_u = torch.norm(u - old_u) / (u.nelement() ** 0.5)
    err_v = torch.norm(v - old_v) / (v.nelement() ** 0.5)
    tol_u = atol + rtol * torch.max(u)
    tol_v = atol + rtol * torch.max(v)
    return bool(err_u < tol_u and err_v < tol_v)


def power_iterate_matrix(W2d, u, v, n_iterations=None, atol=None, rtol=None):
    """Power iteration for a dense (out,in) matrix; returns (u, v, iterations_used).

    Follows ``InducedNormLinear.compute_weight`` (mixed_lipschitz.py:85-123) and
    ``InducedNormConv2d._compute_weight_1x1`` (:276-319) for domain=codomain=2.
    ``rtol`` already carries the ``rtol = ... else atol`` quirk applied by the caller.
    """
    if n_iterations is None and (atol is None or rtol is None):
        raise ValueError('Need one of n_iteration or (atol, rtol).')
    max_itrs = 200 if n_iterations is None else n_iterations
    used = 0
    with torch.no_grad():
        for _ in range(max_itrs):
            old_u, old_v = u, v
            u = _l2n(torch.mv(W2d, v))
            v = _l2n(torch.mv(W2d.t(), u))
            used += 1
            if n_iterations is None and atol is not None and rtol is not None:
                if _tol_reached(u, old_u, v, old_v, atol, rtol):
                    break
    return u, v, used


def power_iterate_conv(W4d, u, v, in_shape, stride, padding, n_iterations=None, atol=None, rtol=None):
    """Power iteration for a k x k conv seen as a linear map on one (c,h,w) image.

    Follows ``InducedNormConv2d._compute_weight_kxk`` (mixed_lipschitz.py:328-376).
    """
    if n_iterations is None and (atol is None or rtol is None):
        raise ValueError('Need one of n_iteration or (atol, rtol).')
    c, h, w = in_shape
    max_itrs = 200 if n_iterations is None else n_iterations
    used = 0
    with torch.no_grad():
        for _ in range(max_itrs):
            old_u, old_v = u, v
            u_s = F.conv2d(v.view(1, c, h, w), W4d, stride=stride, padding=padding, bias=None)
            u = _l2n(u_s.reshape(-1))
            v_s = F.conv_transpose2d(u.view(u_s.shape), W4d, stride=stride, padding=padding, output_padding=0)
            v = _l2n(v_s.reshape(-1))
            used += 1
            if n_iterations is None and atol is not None and rtol is not None:
                if _tol_reached(u, old_u, v, old_v, atol, rtol):
                    break
    return u, v, used


def sigma_matrix(W2d, u, v):
    """sigma = u^T W v  (mixed_lipschitz.py:125, 320)."""
    return torch.dot(u, torch.mv(W2d, v))


def sigma_conv(W4d, u, v, in_shape, stride, padding):
    """sigma = <u, conv(v)>  (mixed_lipschitz.py:378-380)."""
    c, h, w = in_shape
    wv = F.conv2d(v.view(1, c, h, w), W4d, stride=stride, padding=padding, bias=None)
    return torch.dot(u.reshape(-1), wv.reshape(-1))


def soft_rescale(W, sigma, coeff):
    """W / max(1, sigma/coeff); differentiable through sigma (mixed_lipschitz.py:128-131)."""
    factor = torch.max(torch.ones(1), sigma / coeff)
    return W / factor


# --------------------------------------------------------------------------------------
# residual branch container used by the oracle model
# --------------------------------------------------------------------------------------


@dataclass
class OracleLayer:
    """One Lipschitz-constrained affine layer + the activation that PRECEDES it."""
    kind: str                     # 'linear' | 'conv'
    weight: torch.Tensor          # (out,in) or (out,in,k,k); requires_grad for training
    bias: Optional[torch.Tensor]
    u: torch.Tensor
    v: torch.Tensor
    coeff: float
    pre_act: Optional[str] = None  # 'sin' | 'swish' | 'relu' | None — applied to the layer input
    beta: Optional[torch.Tensor] = None  # LipSwish raw beta of that activation
    padding: int = 0
    n_iterations: Optional[int] = None
    atol: Optional[float] = None
    rtol: Optional[float] = None
    spatial: Optional[Sequence[int]] = None  # (h,w) for k x k conv power iteration
    scale: float = 0.0


@dataclass
class OracleBranch:
    """nn.Sequential of [act?] layer act layer ... layer [act?] as built by the train scripts
    (train_toy.py:146-171, train_tabular.py:292-311, implicit_flow.py:359-398,
    train_classification.py:152-167)."""
    layers: List[OracleLayer] = field(default_factory=list)
    post_act: Optional[str] = None   # classifier branches end with ReLU
    post_beta: Optional[torch.Tensor] = None

    def parameters(self):
        ps = []
        for L in self.layers:
            if L.beta is not None:
                ps.append(L.beta)
            ps.append(L.weight)
            if L.bias is not None:
                ps.append(L.bias)
        if self.post_beta is not None:
            ps.append(self.post_beta)
        return ps

    def _act(self, name, x, beta):
        if name is None:
            return x
        if name == 'sin':
            return sin_act(x)
        if name == 'swish':
            return lipswish(x, beta)
        if name == 'relu':
            return torch.relu(x)
        raise ValueError(name)

    def effective_weight(self, L: OracleLayer):
        """``compute_weight(update=False)`` (mixed_lipschitz.py:134-136, 388-391)."""
        if L.kind == 'linear':
            sigma = sigma_matrix(L.weight, L.u, L.v)
        elif L.weight.shape[-1] == 1:
            W2 = L.weight.view(L.weight.shape[0], L.weight.shape[1])
            sigma = sigma_matrix(W2, L.u, L.v)
        else:
            c = L.weight.shape[1]
            sigma = sigma_conv(L.weight, L.u, L.v, (c, L.spatial[0], L.spatial[1]), 1, L.padding)
        L.scale = float(sigma.detach())
        return soft_rescale(L.weight, sigma, L.coeff)

    def __call__(self, x):
        h = x
        for L in self.layers:
            h = self._act(L.pre_act, h, L.beta)
            W = self.effective_weight(L)
            if L.kind == 'linear':
                h = F.linear(h, W, L.bias)
            else:
                h = F.conv2d(h, W, L.bias, 1, L.padding, 1, 1)
        return self._act(self.post_act, h, self.post_beta)

    def update_lipschitz(self, n_iterations=None):
        """``compute_weight(update=True)`` on each layer (train_img.py:786-792)."""
        for L in self.layers:
            n_it = L.n_iterations if n_iterations is None else n_iterations
            atol = L.atol
            rtol = L.rtol if True else None
            # quirk: rtol = self.rtol if rtol is None else atol -> with no override it is self.rtol
            if L.kind == 'linear':
                L.u, L.v, _ = power_iterate_matrix(L.weight.detach(), L.u, L.v, n_it, atol, rtol)
            elif L.weight.shape[-1] == 1:
                W2 = L.weight.detach().view(L.weight.shape[0], L.weight.shape[1])
                L.u, L.v, _ = power_iterate_matrix(W2, L.u, L.v, n_it, atol, rtol)
            else:
                c = L.weight.shape[1]
                L.u, L.v, _ = power_iterate_conv(L.weight.detach(), L.u, L.v, (c, L.spatial[0], L.spatial[1]),
                                                 1, L.padding, n_it, atol, rtol)


# --------------------------------------------------------------------------------------
# Broyden solver  (lib/layers/broyden.py:101-193)
# --------------------------------------------------------------------------------------


def lowrank_left(Us, VTs, x):
    """x^T(-I + U V^T): ``rmatvec`` (broyden.py:101-109). Us (B,d,k), VTs (B,k,d), x (B,d)."""
    if Us.nelement() == 0:
        return -x
    xTU = torch.einsum('bi, bij -> bj', x, Us)
    return -x + torch.einsum('bj, bji -> bi', xTU, VTs)


def lowrank_right(Us, VTs, x):
    """(-I + U V^T) x: ``matvec`` (broyden.py:112-120)."""
    if Us.nelement() == 0:
        return -x
    VTx = torch.einsum('bji, bi -> bj', VTs, x)
    return -x + torch.einsum('bij, bj -> bi', Us, VTx)


def broyden_solve(g_: Callable, x0: torch.Tensor, threshold: int, eps: float):
    """Batched limited-memory good-Broyden root find, line search off (broyden.py:123-193).

    Reproduces: eps*sqrt(B*d) threshold (:131); first update = -g(x0) (:144); batch-global
    norm test (:153,157,163); best-so-far iterate (:159-162); stagnation rule (:165-168);
    protective break (:169-172); rank-1 update with NaN scrub (:174-180); dx=(x+upd)-x (:94,99).
    """
    shape = x0.shape
    x_est = x0.reshape(shape[0], -1)
    bsz, d = x_est.shape
    eps = eps * np.sqrt(np.prod(x_est.shape))

    def g(x):
        return g_(x.view(shape)).reshape(bsz, -1)

    gx = g(x_est)
    nstep = 0
    Us = torch.zeros(bsz, d, threshold).to(x_est)
    VTs = torch.zeros(bsz, threshold, d).to(x_est)
    update = -gx
    new_obj = init_obj = torch.norm(gx).item()
    prot_break = False
    trace = [init_obj]
    lowest = new_obj
    low_x, low_g, low_step = x_est, gx, nstep
    while new_obj >= eps and nstep < threshold:
        x_new = x_est + update                      # line_search(on=False), s = 1 (broyden.py:90-99)
        g_new = g(x_new)
        dx, dg = x_new - x_est, g_new - gx
        x_est, gx = x_new, g_new
        nstep += 1
        new_obj = torch.norm(gx).item()
        trace.append(new_obj)
        if new_obj < lowest:
            low_x, low_g = x_est.clone().detach(), gx.clone().detach()
            lowest, low_step = new_obj, nstep
        if new_obj < eps:
            break
        if new_obj < 3 * eps and nstep == threshold and np.max(trace[-threshold:]) / np.min(trace[-threshold:]) < 1.3:
            break
        if new_obj > init_obj * 1e6:
            prot_break = True
            break
        k = (nstep - 1) % threshold
        pU, pV = Us[:, :, :k], VTs[:, :k]
        vT = lowrank_left(pU, pV, dx)
        u = (dx - lowrank_right(pU, pV, dg)) / torch.einsum('bi, bi -> b', vT, dg)[:, None]
        vT[vT != vT] = 0
        u[u != u] = 0
        VTs[:, k] = vT
        Us[:, :, k] = u
        update = -lowrank_right(Us[:, :, :nstep], VTs[:, :nstep], gx)
    return {"result": low_x.view(shape), "nstep": nstep, "tnstep": nstep, "lowest_step": low_step,
            "diff": torch.norm(low_g).item(), "diff_detail": torch.norm(low_g, dim=1),
            "prot_break": prot_break, "trace": trace, "eps": eps, "threshold": threshold}


def banach_iterate(g: Callable, y: torch.Tensor, threshold: int = 1000, eps: float = 1e-5):
    """``find_fixed_point`` (implicit_block.py:17-28)."""
    x, x_prev = g(y), y
    i = 0
    tol = eps + eps * y.abs()
    while not torch.all((x - x_prev) ** 2 / tol < 1.):
        x, x_prev = g(x), x
        i += 1
        if i > threshold:
            break
    return x


def root_find(net_z, net_x, z0, x, eps, threshold):
    """``RootFind.broyden_find_root`` under no_grad (implicit_block.py:68-91): start from zeros."""
    with torch.no_grad():
        x_embed = net_x(x) + x
        info = broyden_solve(lambda z: x_embed - net_z(z) - z, torch.zeros_like(z0), threshold, eps)
        if info['prot_break']:
            z = banach_iterate(lambda z: x_embed - net_z(z), z0, 1000, eps)
        else:
            z = info['result']
    return z.clone().detach(), info


# --------------------------------------------------------------------------------------
# Russian roulette + estimators  (implicit_block.py:262-350, 418-483)
# --------------------------------------------------------------------------------------


def geometric_1mcdf(p, k, offset):
    """P(N >= k - offset) for the geometric draw (implicit_block.py:461-467)."""
    if k <= offset:
        return 1.
    k = k - offset
    return (1 - p) ** max(k - 1, 0)


def poisson_1mcdf(lamb, k, offset):
    """P(N >= k - offset) for the Poisson draw (implicit_block.py:474-483)."""
    if k <= offset:
        return 1.
    k = k - offset
    s = 1.
    for i in range(1, k):
        s += lamb ** i / math.factorial(i)
    return 1 - np.exp(-lamb) * s


def draw_n(n_dist, n_samples, geom_p=0.5, lamb=2.0):
    """``geometric_sample`` / ``poisson_sample`` on the global NumPy RNG (implicit_block.py:457-471)."""
    if n_dist == 'geometric':
        return np.random.geometric(geom_p, n_samples)
    return np.random.poisson(lamb, n_samples)


def roulette_coefficients(n_draws, n_exact, n_dist, geom_p=0.5, lamb=2.0):
    """n_power_series and coeff_fn(k), k = 1..n  (implicit_block.py:270-289)."""
    n_draws = np.asarray(n_draws)
    n_ps = int(max(n_draws) + n_exact)
    rcdf = (lambda k: geometric_1mcdf(geom_p, k, n_exact)) if n_dist == 'geometric' \
        else (lambda k: poisson_1mcdf(lamb, k, n_exact))
    coeffs = [1 / rcdf(k) * sum(n_draws >= k - n_exact) / len(n_draws) for k in range(1, n_ps + 1)]
    return n_ps, coeffs


def rademacher_like(x):
    """Probe draw on the CPU torch generator (implicit_block.py:297-298)."""
    return torch.distributions.bernoulli.Bernoulli(torch.Tensor([0.5])).sample(x.shape).reshape(x.shape).to(x) * 2 - 1


def logdet_basic(g, x, coeffs, vareps, training):
    """``basic_logdet_estimator`` (implicit_block.py:418-426). coeffs[k-1] = coeff_fn(k)."""
    vjp = vareps
    out = torch.tensor(0.).to(x)
    B = x.shape[0]
    for k in range(1, len(coeffs) + 1):
        vjp = torch.autograd.grad(g, x, vjp, create_graph=training, retain_graph=True)[0]
        tr = torch.sum(vjp.reshape(B, -1) * vareps.reshape(B, -1), 1)
        out = out + (-1) ** (k + 1) / k * coeffs[k - 1] * tr
    return out


def logdet_neumann(g, x, coeffs, vareps, training):
    """``neumann_logdet_estimator`` (implicit_block.py:429-438): surrogate with unbiased gradient."""
    vjp = vareps
    acc = vareps
    B = x.shape[0]
    with torch.no_grad():
        for k in range(1, len(coeffs) + 1):
            vjp = torch.autograd.grad(g, x, vjp, retain_graph=True)[0]
            acc = acc + (-1) ** k * coeffs[k - 1] * vjp
    vjp_jac = torch.autograd.grad(g, x, acc, create_graph=training)[0]
    return torch.sum(vjp_jac.reshape(B, -1) * vareps.reshape(B, -1), 1)


class _MemEffLogDet(torch.autograd.Function):
    """``MemoryEfficientLogDetEstimator`` (implicit_block.py:373-415): backprop-in-forward, backward
    scales the stored grads by grad_out[0] (quirk #13)."""

    @staticmethod
    def forward(ctx, est, net, x, coeffs, vareps, training, *params):
        ctx.training = training
        with torch.enable_grad():
            x = x.detach().requires_grad_(True)
            g = net(x)
            val = est(g, x, coeffs, vareps, training)
            if training:
                gx, *gp = torch.autograd.grad(val.sum(), (x,) + params, retain_graph=True, allow_unused=True)
                if gx is None:
                    gx = torch.zeros_like(x)
                ctx.n = len(params)
                ctx.save_for_backward(gx, *[p if p is not None else torch.zeros(()) for p in gp])
                ctx.none_mask = [p is None for p in gp]
        return val.detach().requires_grad_(val.requires_grad)

    @staticmethod
    def backward(ctx, grad_out):
        if not ctx.training:
            raise ValueError('Provide training=True if using backward.')
        gx, *gp = ctx.saved_tensors
        dL = grad_out[0].detach()
        with torch.no_grad():
            gx = gx * dL
            gp = tuple(None if m else p * dL for p, m in zip(gp, ctx.none_mask))
        return (None, None, gx, None, None, None) + gp


def batch_jacobian(g, x, create_graph=True):
    """(B,d,d) Jacobian by d vjps (implicit_block.py:358-362)."""
    rows = []
    for j in range(g.shape[1]):
        rows.append(torch.autograd.grad(torch.sum(g[:, j]), x, create_graph=create_graph)[0]
                    .view(x.shape[0], 1, x.shape[1]))
    return torch.cat(rows, 1)


def logdet_exact_trace(g, x, coeffs):
    """Power series with the exact trace of the Jacobian powers (implicit_block.py:327-343,
    iresblock.py:147-158): tr(J) + sum_{k>=2} (-1)^(k+1)/k coeff(k) tr(J^k)."""
    J = batch_jacobian(g, x)
    tr = lambda M: M.view(M.shape[0], -1)[:, ::M.shape[1] + 1].sum(1)
    out = tr(J)
    Jk = J
    for k in range(2, len(coeffs) + 1):
        Jk = torch.bmm(J, Jk)
        out = out + (-1) ** (k + 1) / k * coeffs[k - 1] * tr(Jk)
    return out


def logdetgrad(net_x, net_z, z, x, cfg, training, n_draws=None, probes=None):
    """``imBlock._logdetgrad`` (implicit_block.py:245-350) for the branches the configs use:
    brute force (d<=10, eval or brute_force flag) and Hutchinson roulette (basic / Neumann,
    plain / memory-efficient).  Returns (logdet (B,1), n_draws, (vareps_x, vareps_z))."""
    with torch.enable_grad():
        if (cfg['brute_force'] or not training) and (x.ndimension() == 2 and x.shape[1] <= 10):
            x = x.requires_grad_(True)
            z = z.requires_grad_(True)
            Jx = batch_jacobian(x + net_x(x), x)
            Jz = batch_jacobian(z + net_z(z), z)
            return (torch.logdet(Jx) - torch.logdet(Jz)).view(-1, 1), None, None
        if training:
            if cfg.get('n_power_series') is None:
                if n_draws is None:
                    n_draws = draw_n(cfg['n_dist'], cfg['n_samples'], cfg.get('geom_p', 0.5), cfg.get('lamb', 2.0))
                _, coeffs = roulette_coefficients(n_draws, cfg['n_exact_terms'], cfg['n_dist'],
                                                  cfg.get('geom_p', 0.5), cfg.get('lamb', 2.0))
            else:
                coeffs = [1.] * cfg['n_power_series']
        else:
            if n_draws is None:
                n_draws = draw_n(cfg['n_dist'], cfg['n_samples'], cfg.get('geom_p', 0.5), cfg.get('lamb', 2.0))
            _, coeffs = roulette_coefficients(n_draws, cfg['n_exact_terms_test'], cfg['n_dist'],
                                              cfg.get('geom_p', 0.5), cfg.get('lamb', 2.0))
        if cfg.get('exact_trace'):
            x = x.requires_grad_(True)
            z = z.requires_grad_(True)
            ld = logdet_exact_trace(net_x(x), x, coeffs) - logdet_exact_trace(net_z(z), z, coeffs)
            return ld.view(-1, 1), n_draws, None
        if probes is None:
            vx = rademacher_like(x)
            vz = rademacher_like(z)
        else:
            vx, vz = probes
        est = logdet_neumann if (training and cfg['neumann_grad']) else logdet_basic
        if training and cfg['grad_in_forward']:
            ld_x = _MemEffLogDet.apply(est, net_x, x, coeffs, vx, training, *net_x.parameters())
            ld_z = _MemEffLogDet.apply(est, net_z, z, coeffs, vz, training, *net_z.parameters())
        else:
            x = x.requires_grad_(True)
            z = z.requires_grad_(True)
            ld_x = est(net_x(x), x, coeffs, vx, training)
            ld_z = est(net_z(z), z, coeffs, vz, training)
        return (ld_x - ld_z).view(-1, 1), n_draws, (vx, vz)


# --------------------------------------------------------------------------------------
# implicit block forward / backward / inverse  (implicit_block.py:165-243)
# --------------------------------------------------------------------------------------


class _ImplicitBackward(torch.autograd.Function):
    """``imBlock.Backward`` (implicit_block.py:165-217): identity forward, Broyden solve of
    v^T (I + J_z) = grad in backward, then dl_dx = dl_dh (I + J_x)."""

    @staticmethod
    def forward(ctx, net_z, net_x, z, x, eps, threshold, stats):
        ctx.save_for_backward(z, x)
        ctx.cfg = (net_z, net_x, eps, threshold, stats)
        return z

    @staticmethod
    def backward(ctx, grad):
        net_z, net_x, eps, threshold, stats = ctx.cfg
        grad = grad.clone()
        z, x = ctx.saved_tensors
        z = z.clone().detach().requires_grad_()
        x = x.clone().detach().requires_grad_()
        with torch.enable_grad():
            Fz = net_z(z) + z

        def g(v):
            (vJ,) = torch.autograd.grad(Fz, z, v, retain_graph=True)
            return vJ - grad

        info = broyden_solve(g, torch.zeros_like(grad), threshold, eps)
        if stats is not None:
            stats.setdefault('bwd_nstep', []).append(info['nstep'])
            stats.setdefault('bwd_trace', []).append(info['trace'])
        dl_dh = info['result']
        with torch.enable_grad():
            Fx = net_x(x) + x
        (dl_dx,) = torch.autograd.grad(Fx, x, dl_dh)
        return None, None, dl_dh, dl_dx, None, None, None


def imblock_forward(net_x, net_z, x, logpx, cfg, training=True, n_draws=None, probes=None, stats=None):
    """``imBlock.forward`` (implicit_block.py:220-234).  The *_copy nets of the reference hold
    the same weights (load_state_dict, :228-229), so the same branch objects are used here."""
    z0 = x.clone().detach()
    z_star, info = root_find(net_z, net_x, z0, z0, cfg['eps_forward'], cfg['threshold'])
    if stats is not None:
        stats.setdefault('fwd_nstep', []).append(info['nstep'])
        stats.setdefault('fwd_trace', []).append(info['trace'])
    z = net_x(z0) - net_z(z_star.detach()) + z0
    z = _ImplicitBackward.apply(net_z, net_x, z, x, cfg['eps_backward'], cfg['threshold'], stats)
    if logpx is None:
        return z
    ld, n_draws, probes = logdetgrad(net_x, net_z, z, x, cfg, training, n_draws, probes)
    if stats is not None:
        stats.setdefault('n_draws', []).append(None if n_draws is None else np.asarray(n_draws).tolist())
    return z, logpx - ld


def imblock_inverse(net_x, net_z, z, cfg, stats=None):
    """``imBlock.inverse`` without log-det (implicit_block.py:236-240): roles of the nets swapped."""
    x0 = z.clone().detach()
    x, info = root_find(net_x, net_z, x0, z, cfg['eps_sample'], cfg['threshold'])
    if stats is not None:
        stats.setdefault('inv_nstep', []).append(info['nstep'])
    return x


# --------------------------------------------------------------------------------------
# iResBlock  (lib/layers/iresblock.py:54-164, 186-235)
# --------------------------------------------------------------------------------------


class _MemEffLogDetWithOutput(torch.autograd.Function):
    """The iResBlock flavour of ``MemoryEfficientLogDetEstimator`` (iresblock.py:186-235): also returns g
    and, in backward, adds the vjp of grad_g through the kept graph to the scaled stored gradients."""

    @staticmethod
    def forward(ctx, est, net, x, coeffs, vareps, training, *params):
        ctx.training = training
        with torch.enable_grad():
            x = x.detach().requires_grad_(True)
            g = net(x)
            ctx.g, ctx.x = g, x
            val = est(g, x, coeffs, vareps, training)
            if training:
                gx, *gp = torch.autograd.grad(val.sum(), (x,) + params, retain_graph=True, allow_unused=True)
                if gx is None:
                    gx = torch.zeros_like(x)
                ctx.params = params
                ctx.stored = (gx, gp)
        return g.detach().requires_grad_(g.requires_grad), val.detach().requires_grad_(val.requires_grad)

    @staticmethod
    def backward(ctx, grad_g, grad_val):
        if not ctx.training:
            raise ValueError('Provide training=True if using backward.')
        gx, gp = ctx.stored
        with torch.enable_grad():
            dg_x, *dg_p = torch.autograd.grad(ctx.g, [ctx.x] + list(ctx.params), grad_g, allow_unused=True)
        dL = grad_val[0].detach()
        with torch.no_grad():
            gx = gx * dL + dg_x
            out = []
            for d, j in zip(dg_p, gp):
                j = None if j is None else j * dL
                out.append(d if j is None else (j if d is None else d + j))
        return (None, None, gx, None, None, None) + tuple(out)


def ires_logdetgrad(net, x, cfg, training, n_draws=None, probe=None):
    """``iResBlock._logdetgrad`` (iresblock.py:81-164): returns (g(x), logdet (B,1), n_draws, probe).
    2x2 closed form when d == 2 (brute_force flag or eval); eval uses 20 exact terms; Gaussian probe."""
    with torch.enable_grad():
        if (cfg['brute_force'] or not training) and (x.ndimension() == 2 and x.shape[1] == 2):
            x = x.requires_grad_(True)
            g = net(x)
            J = batch_jacobian(g, x)
            det = (J[:, 0, 0] + 1) * (J[:, 1, 1] + 1) - J[:, 0, 1] * J[:, 1, 0]
            return g, torch.log(torch.abs(det)).view(-1, 1), None, None
        if training and cfg.get('n_power_series') is not None:
            coeffs = [1.] * cfg['n_power_series']
        else:
            if n_draws is None:
                n_draws = draw_n(cfg['n_dist'], cfg['n_samples'], cfg.get('geom_p', 0.5), cfg.get('lamb', 2.0))
            _, coeffs = roulette_coefficients(n_draws, cfg['n_exact_terms'] if training else 20, cfg['n_dist'],
                                              cfg.get('geom_p', 0.5), cfg.get('lamb', 2.0))
        if cfg.get('exact_trace'):
            x = x.requires_grad_(True)
            g = net(x)
            return g, logdet_exact_trace(g, x, coeffs).view(-1, 1), n_draws, None
        if probe is None:
            probe = torch.randn_like(x)
        est = logdet_neumann if (training and cfg['neumann_grad']) else logdet_basic
        if training and cfg['grad_in_forward']:
            g, ld = _MemEffLogDetWithOutput.apply(est, net, x, coeffs, probe, training, *net.parameters())
        else:
            x = x.requires_grad_(True)
            g = net(x)
            ld = est(g, x, coeffs, probe, training)
        return g, ld.view(-1, 1), n_draws, probe


def ires_forward(net, x, logpx, cfg, training=True, n_draws=None, probe=None):
    """``iResBlock.forward`` (iresblock.py:54-60)."""
    if logpx is None:
        return x + net(x)
    g, ld, _, _ = ires_logdetgrad(net, x, cfg, training, n_draws, probe)
    return x + g, logpx - ld


def ires_inverse(net, y, atol=1e-5, rtol=1e-5):
    """``iResBlock._inverse_fixed_point`` (iresblock.py:69-79): returns (x, iterations)."""
    x, x_prev = y - net(y), y
    i = 0
    tol = atol + y.abs() * rtol
    while not torch.all((x - x_prev) ** 2 / tol < 1):
        x, x_prev = y - net(x), x
        i += 1
        if i > 1000:
            break
    return x, i


DEFAULT_CFG = dict(brute_force=False, n_dist='geometric', n_samples=1, n_exact_terms=2, n_exact_terms_test=20,
                   n_power_series=None, neumann_grad=True, grad_in_forward=True, eps_forward=1e-6,
                   eps_backward=1e-10, eps_sample=1e-5, threshold=30, geom_p=0.5, lamb=2.0)


# --------------------------------------------------------------------------------------
# glue either side of the path, enough to run a whole flow on the CPU for the baseline
# (act_norm.py:22-62, squeeze.py:32-45, elemwise.py:58-88, train_img.py:135-137,543-549)
# --------------------------------------------------------------------------------------


def actnorm_forward(x, weight, bias, logpx):
    shape = [1, -1] + [1] * (x.dim() - 2)
    y = (x + bias.view(*shape)) * torch.exp(weight.view(*shape))
    ld = weight.view(*shape).expand(*x.size()).contiguous().view(x.size(0), -1).sum(1, keepdim=True)
    return y, logpx - ld


def actnorm_init(x):
    c = x.size(1)
    x_t = x.transpose(0, 1).contiguous().view(c, -1)
    var = torch.max(torch.var(x_t, dim=1), torch.tensor(0.2))
    return (-0.5 * torch.log(var)), (-torch.mean(x_t, dim=1))   # weight, bias


def squeeze2(x):
    b, c, h, w = x.shape
    v = x.reshape(b, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4)
    return v.reshape(b, c * 4, h // 2, w // 2)


def logit_forward(x, alpha, logpx):
    s = alpha + (1 - 2 * alpha) * x
    y = torch.log(s) - torch.log(1 - s)
    ld = (-torch.log(s - s * s) + math.log(1 - 2 * alpha)).view(x.size(0), -1).sum(1, keepdim=True)
    return y, logpx - ld


def bits_per_dim(z, delta_logp, n_dims, nvals=256):
    logpz = (-0.5 * math.log(2 * math.pi) - z.pow(2) / 2).view(z.size(0), -1).sum(1, keepdim=True)
    logpx = logpz - delta_logp - np.log(nvals) * n_dims
    return -torch.mean(logpx) / n_dims / np.log(2)


# --------------------------------------------------------------------------------------
# step tail (lib/optimizers.py:47-107, torch.nn.utils.clip_grad_norm_, lib/utils.py:140-146)
# --------------------------------------------------------------------------------------

def clip_adam_ema_step(params, grads, exp_avg, exp_avg_sq, step, lr, betas, eps, max_norm=None, ema=None,
                       ema_decay=None):
    """One step of train_img.py:652-658 on lists of tensors (updated in place): clip_grad_norm_ over all
    gradients, the vendored Adam (denom = sqrt(v) + eps; step = lr * sqrt(1-b2^t) / (1-b1^t); its weight-decay
    line is a no-op) and ExponentialMovingAverage.apply: the FIRST apply() only copies the (already updated)
    parameters into the shadow (lib/utils.py:140-142), later ones do shadow -= (1 - decay) * (shadow - param)
    (:143-146).  Pinned by tests/golden/step_tail.npz."""
    with torch.no_grad():
        if max_norm is not None:
            total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
            coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
            for g in grads:
                g.mul_(coef)
        b1, b2 = betas
        step_size = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
        for i, (p, g, m, v) in enumerate(zip(params, grads, exp_avg, exp_avg_sq)):
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            p.addcdiv_(m, v.sqrt().add_(eps), value=-step_size)
            if ema is not None:
                if step == 1:
                    ema[i].copy_(p)
                else:
                    ema[i].sub_((1 - ema_decay) * (ema[i] - p))
