"""ResidualFlow: the same multiscale stack built from explicit iResBlocks (y = x + g(x)) instead of imBlocks — API
mirror of lib/resflow.py (ResidualFlow :18-252, StackediResBlocks :255-434; the file is lib/implicit_flow.py with
the block type swapped, so the wiring is shared with implicit_flow.py here).

In the reference this model cannot run: SequentialFlow passes `restore=` to every layer and iResBlock.forward
does not take it.  Here iResBlock.forward accepts and ignores `restore`, so the
constructor / forward / inverse API below is usable."""
from .implicit_flow import ACT_FNS, FCNet, FCWrapper, ImplicitFlow, StackedImplicitBlocks  # noqa: F401

__all__ = ['ResidualFlow', 'StackediResBlocks', 'ACT_FNS']


class StackediResBlocks(StackedImplicitBlocks):
    _implicit = False


class ResidualFlow(ImplicitFlow):
    _stack = StackediResBlocks

    def __init__(self, input_size, n_blocks=[16, 16], intermediate_dim=64, factor_out=True, quadratic=False,
                 init_layer=None, actnorm=False, fc_actnorm=False, batchnorm=False, dropout=0, fc=False, coeff=0.9,
                 vnorms='122f', n_lipschitz_iters=None, sn_atol=None, sn_rtol=None, n_power_series=5,
                 n_dist='geometric', n_samples=1, kernels='3-1-3', activation_fn='elu', fc_end=True, fc_idim=128,
                 n_exact_terms=0, preact=False, neumann_grad=True, grad_in_forward=False, first_resblock=False,
                 learn_p=False, classification=False, classification_hdim=64, n_classes=10, block_type='resblock'):
        if block_type != 'resblock':
            raise NotImplementedError('impflow_b200: block_type=%r (coupling blocks) is a baseline of the reference, '
                                      'outside the ImpFlow hot path' % (block_type,))
        self.block_type = block_type
        super(ResidualFlow, self).__init__(
            input_size, n_blocks=n_blocks, intermediate_dim=intermediate_dim, factor_out=factor_out,
            quadratic=quadratic, init_layer=init_layer, actnorm=actnorm, fc_actnorm=fc_actnorm, batchnorm=batchnorm,
            dropout=dropout, fc=fc, coeff=coeff, vnorms=vnorms, n_lipschitz_iters=n_lipschitz_iters, sn_atol=sn_atol,
            sn_rtol=sn_rtol, n_power_series=n_power_series, n_dist=n_dist, n_samples=n_samples, kernels=kernels,
            activation_fn=activation_fn, fc_end=fc_end, fc_idim=fc_idim, n_exact_terms=n_exact_terms, preact=preact,
            neumann_grad=neumann_grad, grad_in_forward=grad_in_forward, first_resblock=first_resblock,
            learn_p=learn_p, classification=classification, classification_hdim=classification_hdim,
            n_classes=n_classes)
