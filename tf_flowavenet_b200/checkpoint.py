"""TensorFlow checkpoint (TensorBundle, "V2" format) reader / writer without TensorFlow.

The reference saves and restores its variables with ``tf.train.Saver`` (train.py:190,252; synthesize.py:28-34): a checkpoint
``prefix`` is the pair

  prefix.index                   an SSTable (LevelDB table format) mapping   variable name -> BundleEntryProto,
                                 plus the empty key ""                       -> BundleHeaderProto
  prefix.data-00000-of-0000N     the raw little-endian tensor bytes, at (shard_id, offset, size) of each entry

``load_checkpoint(prefix)`` parses both and returns ``{name: numpy array}``; ``flowavenet_variables(prefix)`` strips the
``vocoder/FloWaveNet/`` scope (train.py:53, synthesize.py:11) and drops optimizer slots, giving exactly what
``FloWaveNet.load_variables`` takes.  ``write_checkpoint`` produces the same format (test fixtures, export of trained variables).

Format restated from the published sources (tensorflow/core/util/tensor_bundle/tensor_bundle.cc, tensorflow/core/lib/io/table*.cc
and format.cc = LevelDB's table format, tensorflow/core/protobuf/tensor_bundle.proto); TF itself cannot be installed here, so the
reader is verified against this module's writer and against hand-assembled blocks (prefix-compressed keys, several data blocks,
restart arrays, masked CRC-32C) -- not against a file written by TensorFlow.

Table format.  file = data blocks | metaindex block | index block | footer (48 bytes).  Block = entries + restart array; entry =
varint32 shared | varint32 non_shared | varint32 value_len | key suffix | value; restart array = uint32 offsets + uint32 count.
Every block is followed by a 5-byte trailer: compression type (0 = none, 1 = snappy) and the masked CRC-32C of block + type.
Footer = metaindex handle, index handle (varint64 offset, varint64 size each), zero padding to 40 bytes, magic 0xdb4775248b80fb57.
"""
import os
import struct

import numpy as np

from .dataset import _T as _CRC_TABLES
from .dataset import crc32c as _crc32c_scalar

TABLE_MAGIC = 0xDB4775248B80FB57
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_, 17: np.uint16,
           19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_BFLOAT16 = 14
_DT_OF = {np.dtype(v): k for k, v in _DTYPES.items()}


# ------------------------------------------------------------------------------------------------ CRC-32C of large buffers
_CHUNK = 4096
_T0 = np.array(_CRC_TABLES[0], dtype=np.uint32)
_ZTAB = None


def _zero_shift_tables():
    """Z(r) = CRC register after feeding _CHUNK zero bytes starting from register r (a GF(2)-linear map), as 4 x 256 lookup tables."""
    global _ZTAB
    if _ZTAB is None:
        reg = (np.uint32(1) << np.arange(32, dtype=np.uint32)).astype(np.uint32)   # the 32 basis registers, advanced in lockstep
        for _ in range(_CHUNK):
            reg = (reg >> np.uint32(8)) ^ _T0[reg & np.uint32(0xFF)]
        tabs = np.zeros((4, 256), dtype=np.uint32)
        for byte in range(4):
            for v in range(256):
                acc = np.uint32(0)
                for bit in range(8):
                    if v >> bit & 1:
                        acc ^= reg[8 * byte + bit]
                tabs[byte, v] = acc
        _ZTAB = [[int(x) for x in row] for row in tabs]
    return _ZTAB


def crc32c(data) -> int:
    """CRC-32C (Castagnoli) of bytes / a uint8 array.  Large buffers are cut into 4 KiB chunks whose registers advance in lockstep
    as numpy vectors (the CRC is linear: crc(A || B) = Z_len(B)(crc(A)) xor crc_0(B)); small ones use the scalar routine."""
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8)
    n = buf.size
    if n < (1 << 16):
        return _crc32c_scalar(buf.tobytes())
    nchunk = n // _CHUNK
    body = buf[:nchunk * _CHUNK].reshape(nchunk, _CHUNK)
    reg = np.zeros(nchunk, dtype=np.uint32)
    for j in range(_CHUNK):   # register of every chunk started from zero
        reg = (reg >> np.uint32(8)) ^ _T0[(reg ^ body[:, j]) & np.uint32(0xFF)]
    z0, z1, z2, z3 = _zero_shift_tables()
    r = 0xFFFFFFFF
    for c in reg.tolist():    # chain the chunks: r <- Z(r) xor reg_c
        r = z0[r & 0xFF] ^ z1[(r >> 8) & 0xFF] ^ z2[(r >> 16) & 0xFF] ^ z3[r >> 24] ^ c
    t0 = _CRC_TABLES[0]
    for b in buf[nchunk * _CHUNK:].tolist():
        r = (r >> 8) ^ t0[(r ^ b) & 0xFF]
    return r ^ 0xFFFFFFFF


def masked_crc32c(data) -> int:
    """[TF] lib/hash/crc32c.h Mask(): rotate right by 15 bits and add a constant."""
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ varints / protobuf wire format
def _put_varint(n):
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _get_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf):
    """Yield (field number, wire type, value) of one protobuf message."""
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v, pos = bytes(buf[pos:pos + ln]), pos + ln
        elif wt == 5:
            v, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, v


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_shape(buf):
    dims = []
    for f, _, v in _fields(buf):
        if f == 2:  # TensorShapeProto.Dim
            size = 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    size = _signed(v2)
            dims.append(size)
        elif f == 3 and v:
            raise ValueError("tensor of unknown rank in checkpoint")
    return tuple(dims)


def _parse_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for f, _, v in _fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            e["shape"] = _parse_shape(v)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["slices"] += 1
    return e


def _parse_header(buf):
    h = {"num_shards": 0, "endianness": 0}
    for f, _, v in _fields(buf):
        if f == 1:
            h["num_shards"] = v
        elif f == 2:
            h["endianness"] = v
    return h


# ------------------------------------------------------------------------------------------------ LevelDB table
def _read_block(data, offset, size, verify):
    block, trailer = data[offset:offset + size], data[offset + size:offset + size + 5]
    if len(block) != size or len(trailer) != 5:
        raise ValueError("truncated table block")
    if verify and struct.unpack("<I", trailer[1:])[0] != masked_crc32c(bytes(block) + trailer[:1]):
        raise ValueError("table block checksum mismatch at offset %d" % offset)
    if trailer[0] == 1:
        raise NotImplementedError("snappy-compressed table block (TensorBundle index files are written uncompressed)")
    if trailer[0] != 0:
        raise ValueError("unknown block compression type %d" % trailer[0])
    return block


def _block_entries(block):
    """(key, value) pairs of one block; keys are prefix-compressed against the previous key."""
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of an SSTable file, in key order."""
    with open(path, "rb") as fh:
        data = fh.read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != TABLE_MAGIC:
        raise ValueError("%s is not a TensorBundle index (bad table magic)" % path)
    footer = data[-48:]
    _, p = _get_varint(footer, 0)        # metaindex handle (unused)
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isize, p = _get_varint(footer, p)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, q = _get_varint(handle, 0)
        bsize, q = _get_varint(handle, q)
        out.extend(_block_entries(_read_block(data, boff, bsize, verify)))
    return out


# ------------------------------------------------------------------------------------------------ reader
def load_checkpoint(prefix, verify=True):
    """{variable name: numpy array} of the checkpoint `prefix` (the argument of tf.train.Saver.restore, synthesize.py:34)."""
    entries, header = {}, None
    for key, value in read_table(prefix + ".index", verify):
        if key == b"":
            header = _parse_header(value)
        else:
            entries[key.decode()] = _parse_entry(value)
    if header is None:
        raise ValueError("checkpoint index has no header entry")
    if header["endianness"] != 0:
        raise NotImplementedError("big-endian checkpoint")
    shards, out = {}, {}
    for name, e in entries.items():
        if e["slices"]:
            raise NotImplementedError("partitioned variable '%s' (tensor slices) is not supported" % name)
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, header["num_shards"]), dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if len(raw) != e["size"]:
            raise ValueError("data shard too short for '%s'" % name)
        if verify and e["crc32c"] is not None and masked_crc32c(np.asarray(raw)) != e["crc32c"]:
            raise ValueError("tensor checksum mismatch for '%s'" % name)
        if e["dtype"] == _DT_BFLOAT16:
            arr = (np.frombuffer(raw.tobytes(), dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)
        elif e["dtype"] in _DTYPES:
            arr = np.frombuffer(raw.tobytes(), dtype=_DTYPES[e["dtype"]])
        else:
            raise NotImplementedError("variable '%s' has unsupported dtype enum %d" % (name, e["dtype"]))
        out[name] = arr.reshape(e["shape"]).copy()
    return out


_SLOT_SUFFIXES = ("/Adam", "/Adam_1")


def flowavenet_variables(prefix, scope="vocoder/FloWaveNet", verify=True):
    """The model variables of a reference checkpoint, keyed as FloWaveNet.load_variables expects (names relative to the model scope;
    Adam slots, beta powers and global_step dropped).  Variables are stored in fp32 even in the reference's fp16 mode (utils.py:3-31)."""
    pre = scope.rstrip("/") + "/"
    out = {}
    for name, arr in load_checkpoint(prefix, verify).items():
        if not name.startswith(pre) or name.endswith(_SLOT_SUFFIXES):
            continue
        out[name[len(pre):]] = np.asarray(arr, dtype=np.float32)
    if not out:
        raise ValueError("no variable under scope '%s' in %s" % (scope, prefix))
    return out


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint (synthesize.py:30, train.py:213): the prefix named by the `checkpoint` state file, else the newest index."""
    state = os.path.join(directory, "checkpoint")
    if os.path.exists(state):
        for line in open(state):
            if line.startswith("model_checkpoint_path:"):
                p = line.split(":", 1)[1].strip().strip('"')
                return p if os.path.isabs(p) else os.path.join(directory, p)
    cands = [f[:-6] for f in os.listdir(directory) if f.endswith(".index")]
    if not cands:
        return None
    return os.path.join(directory, max(cands, key=lambda f: os.path.getmtime(os.path.join(directory, f + ".index"))))


# ------------------------------------------------------------------------------------------------ writer
def _ld(field, payload):
    return _put_varint((field << 3) | 2) + _put_varint(len(payload)) + payload


def _vi(field, v):
    return _put_varint(field << 3) + _put_varint(v)


def _entry_proto(dtype_enum, shape, offset, size, crc):
    shp = b"".join(_ld(2, _vi(1, int(d))) for d in shape)
    msg = _vi(1, dtype_enum) + _ld(2, shp)
    if offset:
        msg += _vi(4, offset)
    msg += _vi(5, size) + _put_varint((6 << 3) | 5) + struct.pack("<I", crc)
    return msg


class _BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last, self.interval = bytearray(), [0], 0, b"", restart_interval

    def add(self, key, value):
        shared = 0
        if self.count % self.interval == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def write_table(path, items, block_size=4096):
    """Write sorted (key, value) byte pairs as an uncompressed SSTable."""
    out = bytearray()

    def emit(block):
        off = len(out)
        out.extend(block)
        out.extend(b"\x00" + struct.pack("<I", masked_crc32c(block + b"\x00")))
        return _put_varint(off) + _put_varint(len(block))

    index, bb, last_key = _BlockBuilder(1), _BlockBuilder(), None
    for key, value in items:
        if last_key is not None and key <= last_key:
            raise ValueError("table keys must be strictly increasing")
        bb.add(key, value)
        last_key = key
        if len(bb.buf) >= block_size:
            index.add(key, emit(bb.finish()))
            bb = _BlockBuilder()
    if bb.count:
        index.add(last_key, emit(bb.finish()))
    meta = emit(_BlockBuilder().finish())
    idx = emit(index.finish())
    footer = meta + idx
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    with open(path, "wb") as fh:
        fh.write(out)


def write_checkpoint(prefix, variables):
    """Write {name: array} as a one-shard TensorBundle (what tf.train.Saver.save(sess, prefix) produces, train.py:252)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items, offset = [(b"", _vi(1, 1) + _ld(3, _vi(1, 1)))], 0   # header: num_shards = 1, little endian, version { producer: 1 }
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        for name in sorted(variables, key=lambda s: s.encode()):
            a = np.asarray(variables[name])
            a = np.ascontiguousarray(a) if a.ndim else a        # (ascontiguousarray would promote a scalar to shape (1,))
            if a.dtype not in _DT_OF:
                raise TypeError("unsupported dtype %s for '%s'" % (a.dtype, name))
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            fh.write(raw)
            items.append((name.encode(), _entry_proto(_DT_OF[a.dtype], a.shape, offset, len(raw), masked_crc32c(raw))))
            offset += len(raw)
    write_table(prefix + ".index", items)
    with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as fh:
        fh.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (os.path.basename(prefix), os.path.basename(prefix)))
    return prefix
