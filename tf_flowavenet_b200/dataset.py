"""Input pipeline of the reference (dataset.py, tfrecord.py) without TensorFlow: TFRecord files of tf.train.Example records with
the reference's schema, aligned random crops, per-tower batches in pinned host memory.

Reference behaviour restated here (file:line):
  * schema                     tfrecord.py:17-38  audio FloatList, audio_len Int64List[1], mel_shape Int64List[2], mel FloatList
                                                  (row-major [frames, num_mels]), speaker_id Int64List[1] only if gin_channels > 0
  * padding of short clips     tfrecord.py:41-50  zero-pad audio to max_time_steps and mel by pad // hop_size frames
  * parse + crop               dataset.py:50-85   start ~ U{0 .. frames - max_time_frames - 1}; audio[start*hop : +max_time_steps],
                                                  mel[start : +max_time_frames]   (crop aligned to the hop)
  * shuffle / repeat / batch   dataset.py:22-29   shuffle_and_repeat(buffer 64), batch(batch_size), one batch per tower per step
The record framing ([TF] lib/io/record_writer.cc: u64 length, masked crc32c of the length, payload, masked crc32c of the payload)
and the protobuf wire format of tf.train.Example ([TF] core/example/{example,feature}.proto) are implemented directly, so files
written by the reference's TFRecordCreator can be read and vice versa.
"""
import struct

import numpy as np

# ------------------------------------------------------------------------------------------------ crc32c (Castagnoli), masked
_POLY = 0x82F63B78


def _make_tables():
    t = np.zeros((8, 256), dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (_POLY if (c & 1) else 0)
        t[0, i] = c
    for k in range(1, 8):
        for i in range(256):
            c = int(t[k - 1, i])
            t[k, i] = (c >> 8) ^ int(t[0, c & 0xFF])
    return [[int(v) for v in row] for row in t]


_T = _make_tables()


def crc32c(data: bytes) -> int:
    """Slicing-by-8 CRC-32C of `data`."""
    crc = 0xFFFFFFFF
    n8 = len(data) // 8
    t0, t1, t2, t3, t4, t5, t6, t7 = _T
    if n8:
        for lo, hi in struct.iter_unpack("<II", data[:n8 * 8]):
            lo ^= crc
            crc = (t7[lo & 0xFF] ^ t6[(lo >> 8) & 0xFF] ^ t5[(lo >> 16) & 0xFF] ^ t4[lo >> 24] ^
                   t3[hi & 0xFF] ^ t2[(hi >> 8) & 0xFF] ^ t1[(hi >> 16) & 0xFF] ^ t0[hi >> 24])
    for b in data[n8 * 8:]:
        crc = (crc >> 8) ^ t0[(crc ^ b) & 0xFF]
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    """[TF] lib/hash/crc32c.h Mask(): rotate right by 15 bits and add a constant."""
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ TFRecord framing
class TFRecordWriter:
    def __init__(self, path):
        self._f = open(path, "wb")

    def write(self, record: bytes):
        hdr = struct.pack("<Q", len(record))
        self._f.write(hdr + struct.pack("<I", masked_crc32c(hdr)) + record + struct.pack("<I", masked_crc32c(record)))

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def read_tfrecord(path, verify=True):
    """Yield the raw records of a TFRecord file; IOError on a truncated file or (verify=True) a checksum mismatch."""
    with open(path, "rb") as f:
        while True:
            hdr = f.read(12)
            if not hdr:
                return
            if len(hdr) != 12:
                raise IOError("truncated TFRecord header in %s" % path)
            (n,), (hcrc,) = struct.unpack("<Q", hdr[:8]), struct.unpack("<I", hdr[8:])
            if verify and masked_crc32c(hdr[:8]) != hcrc:
                raise IOError("corrupted TFRecord length in %s" % path)
            body = f.read(n + 4)
            if len(body) != n + 4:
                raise IOError("truncated TFRecord payload in %s" % path)
            if verify and masked_crc32c(body[:n]) != struct.unpack("<I", body[n:])[0]:
                raise IOError("corrupted TFRecord payload in %s" % path)
            yield body[:n]


# ------------------------------------------------------------------------------------------------ tf.train.Example wire format
def _varint(n):
    n &= (1 << 64) - 1   # int64 two's complement
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _ld(field, payload):  # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _feature(value):
    """Feature{bytes_list=1 | float_list=2 | int64_list=3}; FloatList/Int64List{value=1 [packed]}."""
    a = np.asarray(value)
    if a.dtype.kind == "f":
        return _ld(2, _ld(1, a.astype("<f4").tobytes()))
    if a.dtype.kind in "iu":
        return _ld(3, _ld(1, b"".join(_varint(int(v)) for v in a.reshape(-1))))
    raise TypeError("unsupported feature dtype %s" % a.dtype)


def encode_example(features: dict) -> bytes:
    """Example{features=1: Features{feature=1: map<string, Feature>}} -- map entries {key=1, value=2}, keys sorted like TF's
    deterministic serialisation."""
    entries = b"".join(_ld(1, _ld(1, k.encode()) + _ld(2, _feature(v))) for k, v in sorted(features.items()))
    return _ld(1, entries)


def _fields(buf):
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 2:
            n, pos = _read_varint(buf, pos)
            yield field, wt, buf[pos:pos + n]
            pos += n
        elif wt == 0:
            v, pos = _read_varint(buf, pos)
            yield field, wt, v
        elif wt == 5:
            yield field, wt, buf[pos:pos + 4]
            pos += 4
        elif wt == 1:
            yield field, wt, buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)


def _decode_list(kind, buf):
    if kind == 2:    # FloatList: packed (wire type 2) or repeated fixed32 (wire type 5)
        parts = [v for f, wt, v in _fields(buf) if f == 1]
        return np.frombuffer(b"".join(bytes(p) for p in parts), dtype="<f4").copy()
    if kind == 3:    # Int64List: packed varints or repeated varint fields
        out = []
        for f, wt, v in _fields(buf):
            if f != 1:
                continue
            if wt == 0:
                out.append(v)
            else:
                pos = 0
                while pos < len(v):
                    x, pos = _read_varint(v, pos)
                    out.append(x)
        a = np.array(out, dtype=np.uint64).astype(np.int64)   # two's complement
        return a
    if kind == 1:    # BytesList
        return [bytes(v) for f, wt, v in _fields(buf) if f == 1]
    raise ValueError("unknown Feature kind %d" % kind)


def decode_example(record: bytes) -> dict:
    out = {}
    mv = memoryview(record)
    for f, wt, features in _fields(mv):
        if f != 1:
            continue
        for f2, wt2, entry in _fields(features):
            if f2 != 1:
                continue
            key, val = None, None
            for f3, wt3, v in _fields(entry):
                if f3 == 1:
                    key = bytes(v).decode()
                elif f3 == 2:
                    for kind, wt4, lst in _fields(v):
                        val = _decode_list(kind, lst)
            out[key] = val
    return out


# ------------------------------------------------------------------------------------------------ TFRecordCreator (tfrecord.py)
def make_example(audio, mel, speaker_id=None):
    """tfrecord.py:17-38."""
    audio, mel = np.asarray(audio, np.float32), np.asarray(mel, np.float32)
    feats = {"audio": audio, "audio_len": np.array([audio.shape[0]], np.int64), "mel_shape": np.array(mel.shape, np.int64),
             "mel": mel.reshape(-1)}
    if speaker_id is not None:
        feats["speaker_id"] = np.array([int(speaker_id)], np.int64)
    return encode_example(feats)


def adjust_time_resolution(audio, mel, hparams, pad=0.0):
    """tfrecord.py:41-54: zero-pad clips shorter than max_time_steps; len(audio) must then be len(mel) * hop_size."""
    if audio.shape[0] < hparams.max_time_steps:
        audio_pad = hparams.max_time_steps - audio.shape[0]
        mel_pad = audio_pad // hparams.hop_size
        audio = np.pad(audio, (0, audio_pad), mode="constant", constant_values=pad)
        mel = np.pad(mel, ((0, mel_pad), (0, 0)), mode="constant", constant_values=pad)
    assert len(audio) % len(mel) == 0 and len(audio) // len(mel) == hparams.hop_size
    return audio, mel


# ------------------------------------------------------------------------------------------------ Dataset (dataset.py)
class Dataset:
    """dataset.py:8-47.  Iterating yields one list of per-tower batches per step: [(mel [B,frames,num_mels], audio [B,T,1],
    speaker_ids [B] int32 or None)] * num_towers, as torch tensors in pinned host memory when CUDA is present."""

    def __init__(self, tfrecord_path, hparams, num_towers=None, seed=None, buffer_size=64, pin=None):
        self._path, self._hp = tfrecord_path, hparams
        self._max_time_frames = hparams.max_time_steps // hparams.hop_size        # dataset.py:14-15
        self._max_time_steps = self._max_time_frames * hparams.hop_size
        self._towers = int(num_towers if num_towers is not None else getattr(hparams, "num_gpus", 1))
        self._rng = np.random.default_rng(seed)
        self._buffer_size = buffer_size
        self._pin = pin

    def load_sample(self, record):
        """dataset.py:50-85: parse one Example and take the hop-aligned random crop."""
        hp = self._hp
        s = decode_example(record)
        audio_len = int(s["audio_len"][0])
        audio = s["audio"].reshape(audio_len, 1)
        frames, mels = (int(v) for v in s["mel_shape"])
        mel = s["mel"].reshape(frames, mels)
        speaker_id = int(s["speaker_id"][0]) if hp.gin_channels > 0 else 0
        hi = frames - self._max_time_frames
        start = int(self._rng.integers(0, hi)) if hi > 0 else 0   # tf.random.uniform(minval=0, maxval=hi), hi exclusive
        time_start = start * hp.hop_size
        audio = audio[time_start:time_start + self._max_time_steps]
        mel = mel[start:start + self._max_time_frames]
        if mel.shape[1] != hp.num_mels:
            raise ValueError("record has %d mel channels, hparams.num_mels = %d" % (mel.shape[1], hp.num_mels))
        return mel, audio, speaker_id

    def _records(self):
        """shuffle_and_repeat(buffer_size) (dataset.py:24): endless stream, uniform draws from a sliding buffer."""
        buf = []
        while True:
            n = 0
            for rec in read_tfrecord(self._path):
                n += 1
                if len(buf) < self._buffer_size:
                    buf.append(rec)
                    continue
                i = int(self._rng.integers(0, len(buf)))
                out, buf[i] = buf[i], rec
                yield out
            if n == 0:
                raise IOError("empty TFRecord file %s" % self._path)
            if n < self._buffer_size:      # tiny files: the buffer never fills; drain it once per epoch
                order = self._rng.permutation(len(buf))
                for i in order:
                    yield buf[i]
                buf = []

    def __iter__(self):
        import torch
        hp = self._hp
        pin = torch.cuda.is_available() if self._pin is None else self._pin
        recs = self._records()
        while True:
            towers = []
            for _ in range(self._towers):
                samples = [self.load_sample(next(recs)) for _ in range(hp.batch_size)]
                mel = torch.from_numpy(np.stack([s[0] for s in samples]).astype(np.float32))
                audio = torch.from_numpy(np.stack([s[1] for s in samples]).astype(np.float32))
                spk = torch.tensor([s[2] for s in samples], dtype=torch.int32) if hp.gin_channels > 0 else None
                if pin:
                    mel, audio = mel.pin_memory(), audio.pin_memory()
                    spk = spk.pin_memory() if spk is not None else None
                towers.append((mel, audio, spk))
            yield towers
