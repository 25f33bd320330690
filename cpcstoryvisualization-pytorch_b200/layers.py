"""Drop-in for the reference's ``layers.py``: the per-sample dynamic 1-D filter
(reference layers.py:62-80) as ONE batched kernel launch instead of N conv1d calls."""
import torch.nn as nn

from cpcsv_b200 import functions as Fx


class DynamicFilterLayer1D(nn.Module):
    """forward([image (N, 3, L), filters (N, 1, 3, K)]) -> (N, 1, L): cross-correlation of every
    sample with its own filter bank, zero padding ``pad``, summed over the 3 channels."""

    def __init__(self, filter_size, stride=1, pad=0):
        super(DynamicFilterLayer1D, self).__init__()
        self.filter_size, self.stride, self.pad = filter_size, stride, pad
        if stride != 1:
            raise NotImplementedError("the generator only uses stride 1 (reference model.py:310-311)")

    def forward(self, _input, **kwargs):
        image, filters = _input[0], _input[1]
        if self.pad != filters.shape[-1] // 2:
            raise NotImplementedError("only 'same' padding (pad = K // 2) is implemented")
        return Fx.DynamicFilter1dFn.apply(image, filters)
