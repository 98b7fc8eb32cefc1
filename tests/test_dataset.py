"""Input pipeline (dataset.py, tfrecord.py): TFRecord framing, tf.train.Example wire format (cross-checked against
google.protobuf with the real field numbers), the reference's schema and the hop-aligned random crop."""
import os
import struct

import numpy as np
import pytest

from tf_flowavenet_b200 import dataset as D
from tf_flowavenet_b200.hparams import HParams


def test_crc32c_known_answers():
    assert D.crc32c(b"") == 0
    assert D.crc32c(b"123456789") == 0xE3069283
    assert D.crc32c(bytes(32)) == 0x8A9136AA            # RFC 3720 B.4
    assert D.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert D.crc32c(bytes(range(32))) == 0x46DD794E
    data = np.random.default_rng(0).integers(0, 256, 1000, dtype=np.uint8).tobytes()
    crc = 0xFFFFFFFF                                     # bitwise reference
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 if crc & 1 else 0)
    assert D.crc32c(data) == crc ^ 0xFFFFFFFF
    c = D.crc32c(b"abc")
    assert D.masked_crc32c(b"abc") == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _example_classes():
    """tf.train.Example & friends built with google.protobuf from the published field numbers ([TF] core/example/feature.proto)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="fwn_example_test.proto", package="fwn_test", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    for name, typ in (("BytesList", T.TYPE_BYTES), ("FloatList", T.TYPE_FLOAT), ("Int64List", T.TYPE_INT64)):
        m = msg(name)
        f = m.field.add(name="value", number=1, type=typ, label=T.LABEL_REPEATED)
    m = msg("Feature")
    m.oneof_decl.add(name="kind")
    for i, (n, t) in enumerate((("bytes_list", "BytesList"), ("float_list", "FloatList"), ("int64_list", "Int64List")), 1):
        m.field.add(name=n, number=i, type=T.TYPE_MESSAGE, type_name=".fwn_test." + t, label=T.LABEL_OPTIONAL, oneof_index=0)
    m = msg("Features")
    e = m.nested_type.add(name="FeatureEntry")
    e.options.map_entry = True
    e.field.add(name="key", number=1, type=T.TYPE_STRING, label=T.LABEL_OPTIONAL)
    e.field.add(name="value", number=2, type=T.TYPE_MESSAGE, type_name=".fwn_test.Feature", label=T.LABEL_OPTIONAL)
    m.field.add(name="feature", number=1, type=T.TYPE_MESSAGE, type_name=".fwn_test.Features.FeatureEntry", label=T.LABEL_REPEATED)
    m = msg("Example")
    m.field.add(name="features", number=1, type=T.TYPE_MESSAGE, type_name=".fwn_test.Features", label=T.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("fwn_test.Example"))


def test_example_wire_format_interoperates_with_protobuf():
    Example = _example_classes()
    rng = np.random.default_rng(1)
    audio, mel = rng.standard_normal(700).astype(np.float32), rng.random((7, 80)).astype(np.float32)
    # ours -> protobuf
    ex = Example()
    ex.ParseFromString(D.make_example(audio, mel, speaker_id=5))
    f = ex.features.feature
    assert sorted(f) == ["audio", "audio_len", "mel", "mel_shape", "speaker_id"]           # tfrecord.py:27-35
    np.testing.assert_array_equal(np.array(f["audio"].float_list.value, np.float32), audio)
    np.testing.assert_array_equal(np.array(f["mel"].float_list.value, np.float32), mel.reshape(-1))
    assert list(f["audio_len"].int64_list.value) == [700] and list(f["mel_shape"].int64_list.value) == [7, 80]
    assert list(f["speaker_id"].int64_list.value) == [5]
    assert "speaker_id" not in Example.FromString(D.make_example(audio, mel)).features.feature
    # protobuf -> ours (what the reference's TFRecordCreator writes), including a negative int64
    ex = Example()
    ex.features.feature["audio"].float_list.value.extend(audio.tolist())
    ex.features.feature["audio_len"].int64_list.value.append(700)
    ex.features.feature["mel_shape"].int64_list.value.extend([7, 80])
    ex.features.feature["mel"].float_list.value.extend(mel.reshape(-1).tolist())
    ex.features.feature["speaker_id"].int64_list.value.append(-3)
    got = D.decode_example(ex.SerializeToString())
    np.testing.assert_array_equal(got["audio"], audio)
    np.testing.assert_array_equal(got["mel"].reshape(7, 80), mel)
    assert got["audio_len"].tolist() == [700] and got["mel_shape"].tolist() == [7, 80] and got["speaker_id"].tolist() == [-3]


def test_tfrecord_round_trip_and_corruption(tmp_path):
    p = str(tmp_path / "a.tfrecord")
    recs = [b"", b"x", os.urandom(1000), D.make_example(np.arange(10, dtype=np.float32), np.ones((2, 5), np.float32))]
    with D.TFRecordWriter(p) as w:
        for r in recs:
            w.write(r)
    assert list(D.read_tfrecord(p)) == recs
    raw = bytearray(open(p, "rb").read())
    assert struct.unpack("<Q", raw[:8])[0] == 0                      # framing: u64 length first
    raw[12 + 4 + 12] ^= 1                                            # flip a payload byte of the second record
    open(p, "wb").write(raw)
    with pytest.raises(IOError):
        list(D.read_tfrecord(p))
    assert len(list(D.read_tfrecord(p, verify=False))) == len(recs)
    open(p, "wb").write(raw[:-3])
    with pytest.raises(IOError):
        list(D.read_tfrecord(p, verify=False))


def _write_clips(path, hp, lengths, speaker=True):
    with D.TFRecordWriter(path) as w:
        for i, frames in enumerate(lengths):
            audio = (np.arange(frames * hp.hop_size) + 100000 * i).astype(np.float32)      # audio[t] identifies (clip, t)
            mel = np.repeat(np.arange(frames, dtype=np.float32)[:, None], hp.num_mels, 1) + 1000 * i
            audio, mel = D.adjust_time_resolution(audio, mel, hp)
            w.write(D.make_example(audio, mel, i % 7 if speaker else None))


def test_dataset_crop_is_hop_aligned_and_batches_per_tower(tmp_path):
    hp = HParams(hop_size=4, upsample_scales=[2, 2], num_mels=6, max_time_steps=40, batch_size=3, gin_channels=16, n_speakers=7)
    p = str(tmp_path / "train.tfrecord")
    _write_clips(p, hp, [25, 31, 12, 10, 6, 40])                      # clips of 6 and 10 frames are zero-padded to 10 frames
    ds = D.Dataset(p, hp, num_towers=2, seed=0, buffer_size=4, pin=False)
    it = iter(ds)
    seen_clips, starts = set(), set()
    for _ in range(40):
        towers = next(it)
        assert len(towers) == 2
        for mel, audio, spk in towers:
            assert tuple(mel.shape) == (3, 10, 6) and tuple(audio.shape) == (3, 40, 1) and tuple(spk.shape) == (3,)
            for b in range(3):
                clip = int(float(mel[b].max()) // 1000)
                f0 = float(mel[b, 0, 0]) - 1000 * clip
                a0 = float(audio[b, 0, 0]) - 100000 * clip
                assert int(spk[b]) == clip % 7
                if clip == 4:                                         # 6 real frames + padding: only start 0 is possible
                    assert f0 == 0 and float(audio[b, -1, 0]) == 0.0 and float(mel[b, -1, 0]) == 0.0
                assert a0 == f0 * hp.hop_size                         # dataset.py:73-76: time_start = start * hop_size
                np.testing.assert_array_equal(audio[b, :8, 0].numpy() - 100000 * clip, a0 + np.arange(8))
                seen_clips.add(clip)
                starts.add((clip, int(f0)))
    assert seen_clips == {0, 1, 2, 3, 4, 5}
    assert max(s for c, s in starts if c == 5) <= 29 and len({s for c, s in starts if c == 5}) > 5   # start < frames - max_frames


def test_dataset_without_speakers(tmp_path):
    hp = HParams(hop_size=4, upsample_scales=[2, 2], num_mels=6, max_time_steps=16, batch_size=2, gin_channels=-1)
    p = str(tmp_path / "t.tfrecord")
    _write_clips(p, hp, [9, 5], speaker=False)
    mel, audio, spk = next(iter(D.Dataset(p, hp, num_towers=1, seed=1, pin=False)))[0]
    assert spk is None and tuple(mel.shape) == (2, 4, 6) and tuple(audio.shape) == (2, 16, 1)
