import torch

from tensorflow.python.keras.utils import conv_utils
from tensorflow.python.layers import base


class _ConvBase(base.Layer):
    def __init__(self, rank, filters, kernel_size, strides=1, padding="valid", data_format="channels_last", dilation_rate=1,
                 activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None, kernel_regularizer=None,
                 bias_regularizer=None, activity_regularizer=None, kernel_constraint=None, bias_constraint=None,
                 trainable=True, name=None, **kwargs):
        super().__init__(trainable=trainable, name=name, **kwargs)
        self.rank = rank
        self.filters = int(filters)
        self.kernel_size = conv_utils.normalize_tuple(kernel_size, rank)
        self.strides = conv_utils.normalize_tuple(strides, rank)
        self.padding = padding
        self.data_format = data_format
        self.dilation_rate = conv_utils.normalize_tuple(dilation_rate, rank)
        self.activation = activation
        self.use_bias = use_bias
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        self.kernel_regularizer, self.bias_regularizer = kernel_regularizer, bias_regularizer
        self.kernel_constraint, self.bias_constraint = kernel_constraint, bias_constraint


class Conv1D(_ConvBase):
    def __init__(self, filters, kernel_size, **kw):
        super().__init__(1, filters, kernel_size, **kw)

    def call(self, inputs):
        """[TF] keras Conv.call: convolution_op(inputs, kernel) (+ bias) (+ activation)."""
        out = self._convolution_op(inputs, self.kernel)
        if self.use_bias:
            out = out + self.bias
        return self.activation(out) if self.activation is not None else out


class Conv2DTranspose(_ConvBase):
    def __init__(self, filters, kernel_size, strides=(1, 1), **kw):
        super().__init__(2, filters, kernel_size, strides=strides, **kw)

    def call(self, inputs):
        """[TF] keras Conv2DTranspose.call, channels_last, padding 'same': output spatial = input * stride and
        nn.conv2d_transpose == conv2d_backprop_input, i.e. the gradient w.r.t. the input of the SAME-padded
        forward conv (kernel [kh,kw,out_ch,in_ch]).  Computed literally as that gradient via autograd (create_graph keeps it
        differentiable w.r.t. the kernel and the input, for tf.gradients)."""
        assert self.padding == "same" and self.data_format == "channels_last"
        b, h, w, cin = inputs.shape
        kh, kw, cout, _ = self.kernel.shape
        sh, sw = self.strides
        ho, wo = h * sh, w * sw
        ph = max((h - 1) * sh + kh - ho, 0)
        pw = max((w - 1) * sw + kw - wo, 0)
        img = torch.zeros(b, cout, ho, wo, dtype=inputs.dtype, requires_grad=True)
        padded = torch.nn.functional.pad(img, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
        k = self.kernel.as_subclass(torch.Tensor).permute(3, 2, 0, 1)  # [in_ch(out of fwd conv), out_ch, kh, kw]
        y = torch.nn.functional.conv2d(padded, k, stride=(sh, sw))
        assert tuple(y.shape) == (b, cin, h, w), (y.shape, inputs.shape)
        (grad,) = torch.autograd.grad(y, img, grad_outputs=inputs.as_subclass(torch.Tensor).permute(0, 3, 1, 2).contiguous(), create_graph=True)
        import tensorflow as tf
        out = tf.convert_to_tensor(grad.permute(0, 2, 3, 1).contiguous())
        if self.use_bias:
            out = out + self.bias
        return self.activation(out) if self.activation is not None else out
