"""Synthesis caller: the B200 counterpart of the reference's synthesize.py (get_model 10-21, synthesize 23-49, main 51-63).

    python -m tf_flowavenet_b200.synthesize --saved_dir logs/ --mels_dir mels/ --output_dir output/ [--preset hparams8000]
    python -m tf_flowavenet_b200.synthesize --weights model.npz ...

Same wire formats as the reference: every ``*.npy`` in --mels_dir is a float32 mel-spectrogram ``[Tm, num_mels]`` in [0, 1]
(preprocessing.py:68-69); output is a mono wav at hparams.sample_rate named like the mel file.  z ~ N(0,1) * hparams.temp
(synthesize.py:14).  Weights: ``--saved_dir`` is the reference's flag -- a directory of tf.train.Saver checkpoints, of which the
latest is restored (synthesize.py:28-34) through the TF-free TensorBundle reader in ``checkpoint.py`` -- or ``--weights``, either a
checkpoint prefix or an ``.npz`` whose keys are the reference's variable names relative to ``vocoder/FloWaveNet/``.
librosa is not required: wavs are written as 16-bit PCM with the stdlib.
"""
import argparse
import os
import wave

import numpy as np


def write_wav(path, audio, sample_rate):
    """librosa.output.write_wav equivalent for mono float audio in [-1, 1] -> 16-bit PCM."""
    pcm = (np.clip(np.asarray(audio, dtype=np.float64), -1.0, 1.0) * 32767.0).round().astype("<i2")
    with wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(int(sample_rate))
        f.writeframes(pcm.tobytes())


def read_wav(path):
    with wave.open(path, "rb") as f:
        assert f.getnchannels() == 1 and f.getsampwidth() == 2
        return np.frombuffer(f.readframes(f.getnframes()), dtype="<i2").astype(np.float32) / 32767.0, f.getframerate()


def get_model(hparams, weights, device=None):
    """Counterpart of synthesize.get_model + the Saver.restore of synthesize.py:28-34."""
    from . import FloWaveNet, VariableStore
    model = FloWaveNet(hparams, scope="FloWaveNet", variables=VariableStore(), device=device)
    model.load_variables(weights)
    return model


def synthesize_mel(model, hparams, mel, rng):
    """One utterance: mel [Tm, num_mels] -> float32 waveform [Tm * hop]."""
    mel = np.ascontiguousarray(mel, dtype=np.float32)[np.newaxis, ...]           # synthesize.py:44
    hop = int(np.prod(hparams.upsample_scales))
    z = (rng.standard_normal((1, mel.shape[1] * hop, 1)) * hparams.temp).astype(np.float32)  # synthesize.py:14
    return model.reverse_host(z, mel).numpy().reshape(-1)                         # tf.squeeze(predictions)


def load_weights(args):
    """{variable name relative to the model scope: array} from --saved_dir (latest TF checkpoint) or --weights (.npz or checkpoint prefix)."""
    from . import checkpoint
    saved_dir, weights = getattr(args, "saved_dir", None), getattr(args, "weights", None)
    if saved_dir:
        prefix = checkpoint.latest_checkpoint(saved_dir)
        if prefix is None:
            raise FileNotFoundError("no checkpoint in %s" % saved_dir)   # the reference prints and carries on with random weights (synthesize.py:35-37)
        return checkpoint.flowavenet_variables(prefix)
    if weights and weights.endswith(".npz"):
        return dict(np.load(weights))
    if weights:
        return checkpoint.flowavenet_variables(weights)
    raise ValueError("give --saved_dir or --weights")


def synthesize(args, hparams):
    weights = load_weights(args)
    model = get_model(hparams, weights)
    rng = np.random.default_rng(args.seed)
    os.makedirs(args.output_dir, exist_ok=True)
    names = sorted(f for f in os.listdir(args.mels_dir) if f.endswith(".npy"))
    for name in names:
        audio = synthesize_mel(model, hparams, np.load(os.path.join(args.mels_dir, name)), rng)
        write_wav(os.path.join(args.output_dir, name[:-4] + ".wav"), audio, hparams.sample_rate)
    return names


def main():
    from . import hparams as hp_mod
    p = argparse.ArgumentParser()
    p.add_argument("--saved_dir", default=None, help="folder with TF checkpoints (tf.train.Saver); the latest one is restored")
    p.add_argument("--weights", default=None, help=".npz with the reference's variable names, or a TF checkpoint prefix")
    p.add_argument("--mels_dir", default="mels/", help="folder with the mels to synthesize audio from")
    p.add_argument("--output_dir", default="output/", help="folder to contain synthesized audio files")
    p.add_argument("--preset", default="hparams", choices=["hparams", "hparams8000"])
    p.add_argument("--dtype", default="float16", choices=["float32", "float16", "bfloat16"],
                   help="float16 = the reference's mixed dtype (hparams.py:9); float32 = parity mode")
    p.add_argument("--seed", type=int, default=0)
    args = p.parse_args()
    base = getattr(hp_mod, args.preset)
    hparams = hp_mod.HParams(**{**base.values(), "dtype": args.dtype})
    synthesize(args, hparams)


if __name__ == "__main__":
    main()
