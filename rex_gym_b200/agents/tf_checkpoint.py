"""Minimal reader for TensorFlow checkpoint-V2 files (`model.ckpt-N.index` + `.data-00000-of-00001`), enough to load
the PPO policies the reference ships under rex_gym/policies/<task>/<signal>/ without TensorFlow.

Format (public: tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*): the .index file is a LevelDB-style
sorted string table -- data blocks of prefix-compressed (key, value) entries with a restart array, an index block, a
48-byte footer ending in the magic 0xdb4775248b80fb57.  Key "" holds a BundleHeaderProto; every other key is a variable
name whose value is a BundleEntryProto {1: dtype, 2: shape, 3: shard_id, 4: offset, 5: size, 6: crc32c}.  Tensor bytes sit
at [offset, offset+size) of the shard file, little-endian, row-major.
"""
import os
import struct

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}   # tensorflow DataType enum


def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _block(data, offset, size):
    """Return the (key, value) pairs of one table block."""
    raw = data[offset:offset + size]
    ctype = data[offset + size]                       # 1-byte compression type follows the block, then a crc32
    if ctype != 0:
        raise ValueError("compressed checkpoint index blocks are not supported (type %d)" % ctype)
    nrestarts = struct.unpack_from("<I", raw, len(raw) - 4)[0]
    end = len(raw) - 4 - 4 * nrestarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(raw, pos)
        non_shared, pos = _varint(raw, pos)
        vlen, pos = _varint(raw, pos)
        key = key[:shared] + raw[pos:pos + non_shared]
        pos += non_shared
        out.append((key, raw[pos:pos + vlen]))
        pos += vlen
    return out


def _proto_fields(buf):
    """Flat decode of a protobuf message: {field_number: [values]} (varints as int, length-delimited as bytes)."""
    pos, out = 0, {}
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def _shape(buf):
    dims = []
    for d in _proto_fields(buf).get(2, []):           # TensorShapeProto.dim
        dims.append(_proto_fields(d).get(1, [0])[0])  # Dim.size
    return tuple(dims)


def list_variables(prefix):
    """{name: (dtype, shape, shard, offset, size)} of a checkpoint given its prefix (path without .index)."""
    data = open(prefix + ".index", "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != _MAGIC:
        raise ValueError("%s.index is not a TensorFlow checkpoint index" % prefix)
    footer = data[-48:]
    pos = 0
    _, pos = _varint(footer, pos)                     # metaindex handle
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)                  # index handle
    isize, pos = _varint(footer, pos)
    entries = {}
    for _, handle in _block(data, ioff, isize):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, value in _block(data, boff, bsize):
            if key == b"":
                continue                              # BundleHeaderProto
            f = _proto_fields(value)
            dtype = f.get(1, [0])[0]
            entries[key.decode()] = (dtype, _shape(f[2][0]) if 2 in f else (), f.get(3, [0])[0], f.get(4, [0])[0], f.get(5, [0])[0])
    return entries


def load_variables(prefix, names=None):
    """{name: ndarray} for the requested variable names (all float/int variables when None)."""
    ents = list_variables(prefix)
    out, shards = {}, {}
    for name, (dtype, shape, shard, offset, size) in ents.items():
        if names is not None and name not in names:
            continue
        if dtype not in _DTYPES:
            if names is None:
                continue
            raise ValueError("variable %s has unsupported dtype %d" % (name, dtype))
        if shard not in shards:
            nshards = 1
            while not os.path.exists("%s.data-%05d-of-%05d" % (prefix, shard, nshards)) and nshards < 64:
                nshards += 1
            shards[shard] = np.memmap("%s.data-%05d-of-%05d" % (prefix, shard, nshards), dtype=np.uint8, mode="r")
        raw = np.asarray(shards[shard][offset:offset + size])
        out[name] = raw.view(_DTYPES[dtype]).reshape(shape).copy()
    return out


def latest_checkpoint(directory):
    """Prefix of the highest-numbered model.ckpt-N in a policy directory (the reference ships no `checkpoint` file)."""
    best, best_n = None, -1
    for f in os.listdir(directory):
        if f.startswith("model.ckpt-") and f.endswith(".index"):
            n = int(f[len("model.ckpt-"):-len(".index")])
            if n > best_n:
                best, best_n = os.path.join(directory, f[:-len(".index")]), n
    if best is None:
        raise FileNotFoundError("no model.ckpt-*.index under %s" % directory)
    return best
