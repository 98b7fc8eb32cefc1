"""TF-free TensorBundle checkpoint reader / writer (tf.train.Saver files: train.py:190,252; synthesize.py:28-34).

The reader is exercised on (i) files produced by the module's own writer with the reference's variable names, Adam slots and a
global_step; (ii) an index whose BundleEntryProto / BundleHeaderProto messages are serialised by google.protobuf from the published
field numbers ([TF] core/protobuf/tensor_bundle.proto, tensor_shape.proto) and laid out in a hand-assembled table with prefix
compression across several data blocks; (iii) corruption (bad magic, flipped tensor byte, flipped index byte)."""
import os
import struct

import numpy as np
import pytest

from oracle import flowavenet_oracle as O
from tf_flowavenet_b200 import checkpoint as C
from tf_flowavenet_b200 import dataset as D


def test_fast_crc32c_matches_scalar():
    rng = np.random.default_rng(0)
    assert C.crc32c(b"123456789") == 0xE3069283
    for n in (0, 1, 4095, 65535, 65536, 65537, 300001):
        b = rng.integers(0, 256, n, dtype=np.uint8)
        assert C.crc32c(b) == D.crc32c(b.tobytes()), n
        assert C.masked_crc32c(b.tobytes()) == D.masked_crc32c(b.tobytes())


def _reference_style_variables(seed=3):
    hp = O.HP(n_block=2, n_flow=2, n_layer=2, num_mels=8, upsample_scales=(2, 2), gin_channels=4, n_speakers=3)
    rng = np.random.default_rng(seed)
    model = {k: rng.standard_normal(s).astype(np.float32) for k, s in O.param_shapes(hp).items()}
    model["speaker_embeddings"] = rng.standard_normal((3, 4)).astype(np.float32)
    ckpt = {}
    for k, v in model.items():   # what tf.train.Saver(tf.global_variables()) holds under train.py:53's scope
        ckpt["vocoder/FloWaveNet/" + k] = v
        ckpt["vocoder/FloWaveNet/" + k + "/Adam"] = np.zeros_like(v)
        ckpt["vocoder/FloWaveNet/" + k + "/Adam_1"] = np.ones_like(v)
    ckpt["global_step"] = np.array(14000, dtype=np.int64)
    ckpt["beta1_power"] = np.array(0.9, dtype=np.float32)
    ckpt["beta2_power"] = np.array(0.999, dtype=np.float32)
    return hp, model, ckpt


def test_write_read_round_trip(tmp_path):
    hp, model, ckpt = _reference_style_variables()
    prefix = C.write_checkpoint(str(tmp_path / "logs" / "model.ckpt-14000"), ckpt)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    got = C.load_checkpoint(prefix)
    assert sorted(got) == sorted(ckpt)
    for k, v in ckpt.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape
        np.testing.assert_array_equal(got[k], v)
    assert int(got["global_step"]) == 14000
    # the several-hundred keys do not fit one 4 KiB data block: the index block really is walked
    assert os.path.getsize(prefix + ".index") > 3 * 4096
    mv = C.flowavenet_variables(prefix)
    assert sorted(mv) == sorted(model)
    for k in model:
        np.testing.assert_array_equal(mv[k], model[k])
    assert C.latest_checkpoint(str(tmp_path / "logs")) == prefix
    with pytest.raises(ValueError):
        C.flowavenet_variables(prefix, scope="other_scope")


def _bundle_classes():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="fwn_bundle_test.proto", package="fwn_b", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto
    dim = fd.message_type.add(name="Dim")
    dim.field.add(name="size", number=1, type=T.TYPE_INT64, label=T.LABEL_OPTIONAL)
    dim.field.add(name="name", number=2, type=T.TYPE_STRING, label=T.LABEL_OPTIONAL)
    shp = fd.message_type.add(name="TensorShapeProto")
    shp.field.add(name="dim", number=2, type=T.TYPE_MESSAGE, type_name=".fwn_b.Dim", label=T.LABEL_REPEATED)
    shp.field.add(name="unknown_rank", number=3, type=T.TYPE_BOOL, label=T.LABEL_OPTIONAL)
    ver = fd.message_type.add(name="VersionDef")
    ver.field.add(name="producer", number=1, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    hdr = fd.message_type.add(name="BundleHeaderProto")
    hdr.field.add(name="num_shards", number=1, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    hdr.field.add(name="endianness", number=2, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    hdr.field.add(name="version", number=3, type=T.TYPE_MESSAGE, type_name=".fwn_b.VersionDef", label=T.LABEL_OPTIONAL)
    ent = fd.message_type.add(name="BundleEntryProto")
    ent.field.add(name="dtype", number=1, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    ent.field.add(name="shape", number=2, type=T.TYPE_MESSAGE, type_name=".fwn_b.TensorShapeProto", label=T.LABEL_OPTIONAL)
    ent.field.add(name="shard_id", number=3, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    ent.field.add(name="offset", number=4, type=T.TYPE_INT64, label=T.LABEL_OPTIONAL)
    ent.field.add(name="size", number=5, type=T.TYPE_INT64, label=T.LABEL_OPTIONAL)
    ent.field.add(name="crc32c", number=6, type=T.TYPE_FIXED32, label=T.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("fwn_b." + n))
    return get("BundleHeaderProto"), get("BundleEntryProto")


def test_reader_on_protobuf_serialised_two_shard_bundle(tmp_path):
    """Entries serialised by google.protobuf (not by our encoder), tensors spread over two data shards, 2-entry data blocks."""
    Header, Entry = _bundle_classes()
    rng = np.random.default_rng(5)
    tensors = {"vocoder/FloWaveNet/Block_0/Flow_%d/ActNorm/%s" % (j, n): rng.standard_normal((1, 1, 2)).astype(np.float32)
               for j in range(6) for n in ("b", "logs")}
    tensors["vocoder/FloWaveNet/conv2d_transpose/kernel"] = rng.standard_normal((32, 3, 1, 1)).astype(np.float32)
    tensors["vocoder/FloWaveNet/half_precision_probe"] = rng.standard_normal((5,)).astype(np.float16)
    prefix = str(tmp_path / "m.ckpt")
    shards, items = [bytearray(), bytearray()], []
    h = Header(num_shards=2)
    h.version.producer = 1
    items.append((b"", h.SerializeToString()))
    for i, name in enumerate(sorted(tensors)):
        raw = tensors[name].tobytes()
        sid = i % 2
        e = Entry(dtype=1 if tensors[name].dtype == np.float32 else 19, shard_id=sid, offset=len(shards[sid]), size=len(raw),
                  crc32c=D.masked_crc32c(raw))
        for d in tensors[name].shape:
            e.shape.dim.add(size=d)
        shards[sid] += raw
        items.append((name.encode(), e.SerializeToString()))
    for sid in range(2):
        open("%s.data-%05d-of-00002" % (prefix, sid), "wb").write(shards[sid])
    C.write_table(prefix + ".index", items, block_size=64)     # tiny blocks: many data blocks, shared key prefixes inside each
    assert C.read_table(prefix + ".index") == items
    got = C.load_checkpoint(prefix)
    assert sorted(got) == sorted(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype
        np.testing.assert_array_equal(got[k], v)
    # our entry encoder produces what protobuf parses
    e = Entry.FromString(C._entry_proto(1, (3, 256, 512), 1 << 33, 3 * 256 * 512 * 4, 0xDEADBEEF))
    assert (e.dtype, [d.size for d in e.shape.dim], e.offset, e.size, e.crc32c) == (1, [3, 256, 512], 1 << 33, 3 * 256 * 512 * 4, 0xDEADBEEF)


def test_corruption_is_detected(tmp_path):
    _, _, ckpt = _reference_style_variables()
    prefix = C.write_checkpoint(str(tmp_path / "model.ckpt"), ckpt)
    data_path, index_path = prefix + ".data-00000-of-00001", prefix + ".index"
    blob = bytearray(open(data_path, "rb").read())
    blob[100] ^= 0x40
    open(data_path, "wb").write(blob)
    with pytest.raises(ValueError, match="checksum"):
        C.load_checkpoint(prefix)
    assert len(C.load_checkpoint(prefix, verify=False)) == len(ckpt)   # opt-out reads the (corrupt) bytes
    blob[100] ^= 0x40
    open(data_path, "wb").write(blob)
    idx = bytearray(open(index_path, "rb").read())
    idx[50] ^= 1
    open(index_path, "wb").write(idx)
    with pytest.raises(ValueError, match="checksum"):
        C.load_checkpoint(prefix)
    idx[50] ^= 1
    idx[-1] ^= 0xFF
    open(index_path, "wb").write(idx)
    with pytest.raises(ValueError, match="magic"):
        C.load_checkpoint(prefix)
    with pytest.raises(FileNotFoundError):
        C.load_checkpoint(str(tmp_path / "missing.ckpt"))


def test_footer_layout_is_leveldb(tmp_path):
    """48-byte footer: two block handles, zero padding, magic; the index block's keys are the last keys of the data blocks."""
    items = [(("k%04d" % i).encode(), b"v" * (i % 7)) for i in range(300)]
    path = str(tmp_path / "t.sst")
    C.write_table(path, items, block_size=256)
    data = open(path, "rb").read()
    assert struct.unpack("<Q", data[-8:])[0] == 0xDB4775248B80FB57
    assert C.read_table(path) == items
    with pytest.raises(ValueError):
        C.write_table(path, [(b"b", b""), (b"a", b"")])
