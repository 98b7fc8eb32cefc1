import torch


class Convolution:
    """[TF] nn_ops.Convolution for rank-3 channels-last input, VALID padding, stride 1, dilation d:
    y[b,t,o] = sum_k sum_i x[b, t + k*d, i] * w[k,i,o].  Written as an explicit tap sum (einsum)
    on purpose: it shares no code with the oracle's F.conv1d path."""

    def __init__(self, input_shape, filter_shape, dilation_rate, strides, padding, data_format):
        assert padding == "VALID" and tuple(strides) == (1,) and data_format == "NWC", (padding, strides, data_format)
        self.d = int(tuple(dilation_rate)[0])

    def __call__(self, inp, filt):
        k = filt.shape[0]
        t_out = inp.shape[1] - self.d * (k - 1)
        out = None
        for j in range(k):
            term = torch.einsum("bti,io->bto", inp[:, j * self.d: j * self.d + t_out], filt[j])
            out = term if out is None else out + term
        return out
