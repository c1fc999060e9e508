"""SequentialFlow — mirror of lib/layers/container.py:4-30."""
import torch.nn as nn

__all__ = ['SequentialFlow', 'Inverse']


class SequentialFlow(nn.Module):
    """nn.Sequential for flows: threads (x, logpx) and the `restore` flag through the chain."""

    def __init__(self, layersList):
        super(SequentialFlow, self).__init__()
        self.chain = nn.ModuleList(layersList)

    def forward(self, x, logpx=None, restore=False):
        if logpx is None:
            for layer in self.chain:
                x = layer(x, restore=restore)
            return x
        for layer in self.chain:
            x, logpx = layer(x, logpx, restore=restore)
        return x, logpx

    def inverse(self, y, logpy=None):
        if logpy is None:
            for layer in reversed(self.chain):
                y = layer.inverse(y)
            return y
        for layer in reversed(self.chain):
            y, logpy = layer.inverse(y, logpy)
        return y, logpy


class Inverse(nn.Module):

    def __init__(self, flow):
        super(Inverse, self).__init__()
        self.flow = flow

    def forward(self, x, logpx=None):
        return self.flow.inverse(x, logpx)

    def inverse(self, y, logpy=None):
        return self.flow.forward(y, logpy)
