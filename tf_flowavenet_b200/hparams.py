"""Hyper-parameters of the path: the fields FloWaveNet.__init__ reads (reference hparams.py:6-50, model.py:288-314)
as a plain object, plus the two shipped presets.  ``dtype`` selects the numeric mode:
'float32' = fp32 parity mode; 'float16' = fp16 operands with fp32 accumulate, the reference's own mixed dtype (hparams.py:9-10:
fp16 compute, fp32 master weights, loss scale 64 for training); 'bfloat16' (alias 'mixed') = bf16 operands with fp32 accumulate."""


class HParams:
    def __init__(self, **kw):
        self.__dict__.update(dict(
            dtype="float32", num_mels=80, hop_size=256, sample_rate=22050, max_time_steps=6400, batch_size=8,
            gin_channels=-1, n_speakers=7, causal=False, n_block=8, n_flow=6, n_layer=2, affine=True, causality=False,
            temp=0.7, upsample_scales=[16, 16], filter_size=256))
        self.__dict__.update(kw)

    def values(self):
        return dict(self.__dict__)

    def __repr__(self):
        return "HParams(%s)" % ", ".join("%s=%r" % kv for kv in sorted(self.__dict__.items()))


# reference hparams.py (22.05 kHz LJSpeech) and hparams8000.py (8 kHz; differs in hop/sample_rate/n_block/scales)
hparams = HParams()
hparams8000 = HParams(hop_size=96, sample_rate=8000, max_time_steps=2320, n_block=5, upsample_scales=[8, 12])
