"""Fused, graph-free evaluation of a residual branch for the no-grad hot loops.

The three loops that dominate an ImpFlow step never need an autograd graph:
  * the (nstep+1) branch evaluations of a forward / inverse Broyden solve (implicit_block.py:68-80),
  * the vjps of the implicit-differentiation solve in backward (implicit_block.py:199-207),
  * the n_power_series vjps of the Neumann estimator (implicit_block.py:432-435).
A BranchProgram compiles an nn.Sequential of [act] layer act layer ... [act] (InducedNormLinear /
InducedNormConv2d 1x1 / 3x3 + Sin / Swish / ReLU — every branch the shipped configs build) into a
short list of kernel launches per evaluation:
  forward : [act] -> (im2col) GEMM{bias, act, hi/lo split} -> ... -> GEMM (col2im{bias})
  vjp     : GEMM^T{act', hi/lo split} ... with the pre-activations saved by forward(save=True)
with the spectrally rescaled weights (compute_weight(update=False), hoisted out of the loops —
they are constant within a solve), their conv-as-GEMM re-layouts and tf32 hi/lo planes cached until
a parameter or u/v buffer changes version.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi, ops
from .layers.base.activations import ReLU, Sin, Swish
from .layers.base.mixed_lipschitz import InducedNormConv2d, InducedNormLinear

__all__ = ['BranchProgram', 'compile_branch']


def _round_up(n, m):
    return (n + m - 1) // m * m


class _Act(object):
    def __init__(self, kind, module=None):
        self.kind = kind
        self.module = module       # Swish module (for beta) or None
        self._beta = None          # softplus(beta) device scalar, refreshed by BranchProgram._prep

    def refresh(self):
        if self.kind == ops.ACT_LIPSWISH:
            self._beta = F.softplus(self.module.beta.detach())

    def beta_sp(self):
        if self.kind != ops.ACT_LIPSWISH:
            return None
        if self._beta is None:
            self.refresh()
        return self._beta


class _Weights(object):
    """Effective weights of one layer in the layouts the fused path needs (fp32 + optional planes)."""
    __slots__ = ('fwd', 'fwd_split', 'bwd', 'bwd_split', 'bias', 'kind', 'cin', 'cout', 'fwd_k', 'bwd_k', 'a_type')


def _act_of(m):
    if isinstance(m, Sin):
        return _Act(ops.ACT_SIN)
    if isinstance(m, Swish):
        return _Act(ops.ACT_LIPSWISH, m)
    if isinstance(m, (ReLU, nn.ReLU)):
        return _Act(ops.ACT_RELU)
    return None


def compile_branch(nnet):
    """BranchProgram for `nnet`, or None if it contains a module this path does not know."""
    flatten = False
    seq = nnet
    if hasattr(nnet, 'nnet') and isinstance(getattr(nnet, 'nnet'), nn.Sequential) and hasattr(nnet, 'input_shape'):
        seq, flatten = nnet.nnet, True            # FCNet wrapper (implicit_flow.py:437-474)
    if not isinstance(seq, nn.Sequential):
        return None
    stages, pending = [], None
    for m in seq:
        a = _act_of(m)
        if a is not None:
            if pending is not None:
                return None
            pending = a
        elif isinstance(m, (InducedNormLinear, InducedNormConv2d)):
            stages.append((pending, m))
            pending = None
        else:
            return None
    if not stages:
        return None
    kinds = {isinstance(m, InducedNormLinear) for _, m in stages}
    if len(kinds) != 1:
        return None
    return BranchProgram(stages, pending, flatten)


class BranchProgram(object):

    def __init__(self, stages, post_act, flatten):
        self.stages = stages            # [(pre_act or None, layer module)]
        self.post_act = post_act
        self.flatten = flatten
        self.is_linear = isinstance(stages[0][1], InducedNormLinear)
        self._key = None
        self._weights = None
        self._saved = None

    # ---------------------------------------------------------------- weights
    def _version_key(self):
        key = []
        for act, m in self.stages:
            key += [m.weight._version, m.u._version, m.v._version, m.weight.data_ptr(),
                    (m.bias._version if m.bias is not None else -1)]
        for act in [a for a, _ in self.stages] + [self.post_act]:
            if act is not None and act.module is not None:
                key += [act.module.beta._version, act.module.beta.data_ptr()]
        return tuple(key)

    def _use_tc(self, M, N, K):
        mode = ops.get_gemm_backend()
        if mode == 'simt':
            return False
        if mode == 'tc':
            return True
        return M >= 64 and N >= 8 and K >= 32

    def _prep(self, M, meta=None):
        """(Re)build the cached effective weights; M = rows of the activation matrices."""
        key = (self._version_key(), M, ops.get_gemm_backend())
        if key == self._key:
            return self._weights
        ws = []
        with torch.no_grad():
            for act in [a for a, _ in self.stages] + [self.post_act]:
                if act is not None:
                    act.refresh()
            for act, m in self.stages:
                w = _Weights()
                if isinstance(m, InducedNormConv2d) and not m.initialized and meta is not None:
                    # first use: record the spatial dims like InducedNormConv2d.forward does
                    m.spatial_dims.copy_(torch.tensor([float(meta[1][1]), float(meta[1][2])]).to(m.spatial_dims))
                    m._hw = None
                W = m.compute_weight(update=False).detach()
                w.bias = m.bias.detach() if m.bias is not None else None
                if isinstance(m, InducedNormLinear) or m.kernel_size == (1, 1):
                    cout, cin = W.shape[0], W.shape[1]
                    W2 = W.reshape(cout, cin)
                    w.kind, w.cin, w.cout, w.a_type = 'mm', cin, cout, True
                    fwd, bwd = W2, W2.t()
                else:
                    cout, cin = W.shape[0], W.shape[1]
                    w.kind, w.cin, w.cout = 'c3', cin, cout
                    w.a_type = cin <= cout
                    if w.a_type:     # forward: im2col + GEMM(K=9cin); transpose: GEMM(N=9cin) + col2im
                        fwd = W.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
                        bwd = W.permute(2, 3, 1, 0).reshape(9 * cin, cout)
                    else:            # forward: GEMM(N=9cout) + col2im; transpose: im2col + GEMM(K=9cout)
                        fwd = W.flip(2, 3).permute(2, 3, 0, 1).reshape(9 * cout, cin)
                        bwd = W.flip(2, 3).permute(1, 2, 3, 0).reshape(cin, 9 * cout)
                w.fwd, w.fwd_k, w.fwd_split = self._finish(fwd, M)
                w.bwd, w.bwd_k, w.bwd_split = self._finish(bwd, M)
                ws.append(w)
        self._key, self._weights = key, ws
        return ws

    def _finish(self, Wm, M):
        """Contiguous (N, Kpad) weight matrix, its padded K and (if the tcgen05 path applies) planes."""
        N, K = Wm.shape
        if self._use_tc(M, N, _round_up(K, 32)):
            Kp = _round_up(K, 32)
            if Kp != K:
                Wp = Wm.new_zeros(N, Kp)
                Wp[:, :K] = Wm
            else:
                Wp = Wm.contiguous()
            return Wp, Kp, ops.split_tf32(Wp)
        return Wm.contiguous(), K, None

    # ---------------------------------------------------------------- plumbing
    def _to_rows(self, x):
        """Module-level tensor -> (rows matrix, meta)."""
        if self.is_linear:
            x2 = x.reshape(x.shape[0], -1) if self.flatten else x.reshape(-1, x.shape[-1])
            return x2.contiguous(), ('lin', tuple(x.shape))
        B, C, H, W = x.shape
        rows = x.permute(0, 2, 3, 1).contiguous().view(B * H * W, C)
        return rows, ('conv', (B, H, W))

    def _from_rows(self, y, meta):
        if meta[0] == 'lin':
            shape = meta[1]
            return y.view(*shape[:-1], y.shape[-1]) if not self.flatten else y.view(*shape)
        B, H, W = meta[1]
        return y.view(B, H, W, y.shape[-1]).permute(0, 3, 1, 2)

    def _gemm(self, A, A_split, Wm, Wk, W_split, bias, act, want_pre, want_act, dmul_pre, want_split):
        """One fused GEMM launch; A is (M, Wk) fp32 and/or its planes."""
        lib = _cabi.load()
        M = (A if A is not None else A_split[0]).shape[0]
        N = Wm.shape[0]
        dev = Wm.device
        kind = act.kind if act is not None else ops.ACT_NONE
        beta = act.beta_sp() if act is not None else None
        if W_split is None and not want_act:
            want_pre = True            # the CUDA-core kernel has no plane outputs
        pre = torch.empty(M, N, device=dev, dtype=torch.float32) if want_pre else None
        out_act = torch.empty(M, N, device=dev, dtype=torch.float32) if want_act else None
        if W_split is not None:
            if A_split is None:
                A_split = ops.split_tf32(A)
            sh = torch.empty(M, N, device=dev, dtype=torch.float32) if want_split else None
            sl = torch.empty(M, N, device=dev, dtype=torch.float32) if want_split else None
            if ops.GEMM_PROFILE['on']:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
            _cabi.check(lib.impflow_gemm_nt_tc(
                _cabi.ptr(A_split[0]), _cabi.ptr(A_split[1]), Wk, _cabi.ptr(W_split[0]), _cabi.ptr(W_split[1]), Wk,
                _cabi.ptr(bias, 'bias', True), _cabi.ptr(pre, 'pre', True), _cabi.ptr(out_act, 'act', True),
                _cabi.ptr(dmul_pre, 'dmul', True), _cabi.ptr(sh, 'sh', True), _cabi.ptr(sl, 'sl', True), N, M, N, Wk,
                kind, _cabi.ptr(beta, 'beta', True), None, _cabi.stream()), 'gemm_nt_tc')
            if ops.GEMM_PROFILE['on']:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                ops.GEMM_PROFILE['events'].append((e0, e1, 2.0 * M * N * Wk))
            return pre, out_act, ((sh, sl) if want_split else None)
        if A is None:
            A = ops.lincomb3(A_split[0], 1.0, A_split[1], 1.0)
        _cabi.check(lib.impflow_gemm_nt(
            _cabi.ptr(A), Wk, _cabi.ptr(Wm), Wk, _cabi.ptr(bias, 'bias', True), _cabi.ptr(pre, 'pre', True),
            _cabi.ptr(out_act, 'act', True), _cabi.ptr(dmul_pre, 'dmul', True), N, M, N, Wk, kind,
            _cabi.ptr(beta, 'beta', True), _cabi.stream()), 'gemm_nt')
        return pre, out_act, None

    @staticmethod
    def _pad_cols(A, Kp):
        if A.shape[1] == Kp:
            return A
        out = A.new_zeros(A.shape[0], Kp)
        out[:, :A.shape[1]] = A
        return out

    # ---------------------------------------------------------------- forward
    def forward(self, x, save=False):
        """nnet(x) without a graph.  save=True keeps the pre-activations for vjp()."""
        rows, meta = self._to_rows(x)
        M = rows.shape[0]
        ws = self._prep(M, meta)
        saved = []                    # saved[i] = input of the activation in front of layer i (or None)
        n = len(self.stages)
        h, h_split = rows, None
        act0 = self.stages[0][0]
        if act0 is not None:
            saved.append(rows if save else None)
            h = ops.act_mul(rows, None, act0.kind, 0, act0.beta_sp())
        else:
            saved.append(None)
        out = None
        for i, (_, m) in enumerate(self.stages):
            w = ws[i]
            last = i == n - 1
            nxt = self.post_act if last else self.stages[i + 1][0]
            need_pre = (nxt is None) or (save and nxt is not None)
            if w.kind == 'mm' or w.a_type:
                if w.kind == 'c3':
                    B, H, Wd = meta[1]
                    A = ops.im2col3x3(h.view(B, H, Wd, w.cin), ld=w.fwd_k)
                    A_split = None
                else:
                    A_split = h_split if (h_split is not None and h_split[0].shape[1] == w.fwd_k) else None
                    A = self._pad_cols(h, w.fwd_k) if (A_split is None or w.fwd_split is None) else None
                nxt_tc = (not last) and ws[i + 1].fwd_split is not None and \
                    (ws[i + 1].kind == 'mm' or not ws[i + 1].a_type) and ws[i + 1].fwd_k == w.cout
                emit_split = nxt_tc and w.fwd_split is not None
                want_act = nxt is not None and (last or not emit_split or ws[i + 1].fwd_split is None)
                pre, a_out, split = self._gemm(A, A_split, w.fwd, w.fwd_k, w.fwd_split, w.bias, nxt,
                                               want_pre=need_pre, want_act=want_act, dmul_pre=None,
                                               want_split=emit_split)
                if nxt_tc and not emit_split:     # this GEMM ran on CUDA cores: planes by the split kernel
                    split = ops.split_tf32(a_out if nxt is not None else pre)
            else:   # narrow-output 3x3: GEMM to (M, 9*cout) then col2im with the fused epilogue
                B, H, Wd = meta[1]
                A_split = h_split if (h_split is not None and h_split[0].shape[1] == w.fwd_k) else None
                A = self._pad_cols(h, w.fwd_k) if (A_split is None or w.fwd_split is None) else None
                Y, _, _ = self._gemm(A, A_split, w.fwd, w.fwd_k, w.fwd_split, None, None, True, False, None, False)
                pre, a_out = ops.col2im3x3(Y, B, H, Wd, w.cout, w.bias, nxt.kind if nxt is not None else ops.ACT_NONE,
                                           nxt.beta_sp() if nxt is not None else None,
                                           want_pre=need_pre, want_act=nxt is not None)
                pre = pre.view(M, w.cout) if pre is not None else None
                a_out = a_out.view(M, w.cout) if a_out is not None else None
                split = None
            if nxt is not None:
                saved.append(pre if save else None)
            else:
                saved.append(None)
            if last:
                out = a_out if nxt is not None else pre
            else:
                h, h_split = (a_out if nxt is not None else pre), split
                if h is None and h_split is None:
                    raise RuntimeError('BranchProgram: lost the activation of layer %d' % i)
                if h is None and not (ws[i + 1].fwd_split is not None and h_split[0].shape[1] == ws[i + 1].fwd_k):
                    h = ops.lincomb3(h_split[0], 1.0, h_split[1], 1.0)
        if save:
            self._saved = (saved, meta, M)
        return self._from_rows(out, meta)

    # ---------------------------------------------------------------- vjp
    def vjp(self, v):
        """v^T J at the point of the last forward(save=True)."""
        if self._saved is None:
            raise RuntimeError('BranchProgram.vjp: call forward(save=True) first')
        saved, meta, M = self._saved
        ws = self._prep(M)
        t, _ = self._to_rows(v)
        t_split = None
        n = len(self.stages)
        if self.post_act is not None:
            t = ops.act_mul(saved[n], t, self.post_act.kind, 1, self.post_act.beta_sp())
        for i in range(n - 1, -1, -1):
            w = ws[i]
            act = self.stages[i][0]            # activation in front of layer i: multiply by act'(saved[i])
            dm = saved[i] if act is not None else None
            if w.kind == 'mm' or not w.a_type:
                # transpose is "wide K": plain GEMM (mm) or im2col + GEMM (narrow-output conv)
                if w.kind == 'c3':
                    B, H, Wd = meta[1]
                    A = ops.im2col3x3(t.view(B, H, Wd, w.cout), ld=w.bwd_k)
                    A_split = None
                else:
                    A_split = t_split if (t_split is not None and t_split[0].shape[1] == w.bwd_k) else None
                    A = self._pad_cols(t, w.bwd_k) if (A_split is None or w.bwd_split is None) else None
                nxt_tc = i > 0 and ws[i - 1].bwd_split is not None and \
                    (ws[i - 1].kind == 'mm' or ws[i - 1].a_type) and ws[i - 1].bwd_k == w.cin
                emit_split = nxt_tc and w.bwd_split is not None
                pre, _, split = self._gemm(A, A_split, w.bwd, w.bwd_k, w.bwd_split, None, act, want_pre=not emit_split,
                                           want_act=False, dmul_pre=dm, want_split=emit_split)
                if nxt_tc and not emit_split:
                    split = ops.split_tf32(pre)
                t, t_split = pre, split
            else:
                B, H, Wd = meta[1]
                A_split = t_split if (t_split is not None and t_split[0].shape[1] == w.bwd_k) else None
                A = self._pad_cols(t, w.bwd_k) if (A_split is None or w.bwd_split is None) else None
                Y, _, _ = self._gemm(A, A_split, w.bwd, w.bwd_k, w.bwd_split, None, None, True, False, None, False)
                pre, _ = ops.col2im3x3(Y, B, H, Wd, w.cin, None, act.kind if act is not None else ops.ACT_NONE,
                                       act.beta_sp() if act is not None else None, want_pre=True, want_act=False,
                                       dmul_pre=dm.view(B, H, Wd, w.cin) if dm is not None else None)
                t, t_split = pre.view(M, w.cin), None
        if t is None:
            t = ops.lincomb3(t_split[0], 1.0, t_split[1], 1.0)
        return self._from_rows(t, meta)
