"""B200-native FloWaveNet flow pass: drop-in for the model.py / modules.py API of ryhorv/tf-flowavenet.

Host code is Python; all arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
libflowavenet_b200.so (include/flowavenet_b200.h).  PyTorch is used only to own device memory and streams.
"""
from .hparams import HParams, hparams, hparams8000  # noqa: F401
from .model import ActNorm, AffineCoupling, Block, Flow, FloWaveNet, change_order  # noqa: F401
from .modules import Conv, ResBlock, WaveNet, ZeroConv1d  # noqa: F401
from .variables import VariableStore, variable_scope  # noqa: F401
