"""TensorFlow tensor-bundle (checkpoint v2) reader without TensorFlow.

The reference restores its weights with ``tf.train.Saver.restore`` (reference
src/process_functional.py:32, :43).  The only trained artefact it ships is
``data/tensorboard_log/model_epoch2000.ckpt.{index,data-00000-of-00001}``.  This module reads
that format directly: the ``.index`` file is a LevelDB-style SSTable (prefix-compressed keys,
uncompressed blocks) whose values are ``BundleEntryProto`` messages; the ``.data`` shard holds
raw little-endian tensors.  Every tensor's masked CRC32C is verified.

Layout facts decoded in SURVEY.md Appendix B.
"""
import os
import struct

import numpy as np

_TABLE_MAGIC = 0xdb4775248b80fb57
_DT_FLOAT = 1


def _varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _parse_block(data, offset, size):
    """Yield (key, value) pairs of one SSTable block (restart array ignored, keys rebuilt)."""
    block = data[offset:offset + size]
    block_type = data[offset + size]
    if block_type != 0:
        raise ValueError("compressed SSTable blocks are not supported (type %d)" % block_type)
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos = 0
    key = b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        value_len, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        value = bytes(block[pos:pos + value_len])
        pos += value_len
        yield key, value


def _parse_proto(buf):
    """Minimal protobuf wire decoder -> list of (field, wire_type, value)."""
    out = []
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((field, wt, v))
    return out


def _parse_entry(value):
    """BundleEntryProto: 1 dtype, 2 shape, 3 shard_id, 4 offset, 5 size, 6 crc32c (fixed32)."""
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None)
    for field, _, v in _parse_proto(value):
        if field == 1:
            e["dtype"] = v
        elif field == 2:
            dims = []
            for f2, _, v2 in _parse_proto(v):
                if f2 == 2:          # TensorShapeProto.dim
                    size = 0
                    for f3, _, v3 in _parse_proto(v2):
                        if f3 == 1:
                            size = v3
                    dims.append(size)
            e["shape"] = dims
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc32c"] = v
    return e


_CRC_TABLE = None


def crc32c(data):
    """CRC-32C (Castagnoli), table driven; vectorised 8 KB chunks are unnecessary at 1.2 MB."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tbl = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tbl.append(c)
        _CRC_TABLE = tbl
    tbl = _CRC_TABLE
    crc = 0xFFFFFFFF
    for b in data:
        crc = tbl[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data):
    crc = crc32c(data)
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def read_index(prefix):
    """Return {tensor_name: entry dict} for a checkpoint prefix (path without .index)."""
    with open(prefix + ".index", "rb") as f:
        data = f.read()
    if len(data) < 48:
        raise ValueError("index file too short")
    footer = data[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != _TABLE_MAGIC:
        raise ValueError("bad SSTable magic in %s.index" % prefix)
    pos = 0
    _, pos = _varint(footer, pos)       # metaindex handle
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries = {}
    for _, handle in _parse_block(data, idx_off, idx_size):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, value in _parse_block(data, boff, bsize):
            if key == b"":
                continue                 # BundleHeaderProto
            entries[key.decode("utf-8")] = _parse_entry(value)
    return entries


def load_checkpoint(prefix, verify_crc=True, skip_slots=True):
    """Read every float32 tensor of a TF bundle -> {name: np.ndarray}.

    ``skip_slots`` drops optimizer slots (``.../Momentum``), which inference does not need.
    """
    entries = read_index(prefix)
    out = {}
    shards = {}
    for name, e in sorted(entries.items()):
        if skip_slots and name.rsplit("/", 1)[-1] not in ("weights", "biases"):
            continue
        if e["dtype"] != _DT_FLOAT:
            continue
        sid = e["shard_id"]
        if sid not in shards:
            # single-shard bundles only need 00000-of-00001; general form kept for clarity
            cands = [p for p in os.listdir(os.path.dirname(prefix) or ".")
                     if p.startswith(os.path.basename(prefix) + ".data-%05d-of-" % sid)]
            if not cands:
                raise FileNotFoundError("data shard %d of %s not found" % (sid, prefix))
            with open(os.path.join(os.path.dirname(prefix) or ".", cands[0]), "rb") as f:
                shards[sid] = f.read()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if len(raw) != e["size"]:
            raise ValueError("tensor %s truncated" % name)
        if verify_crc and e["crc32c"] is not None and masked_crc32c(raw) != e["crc32c"]:
            raise ValueError("CRC32C mismatch for tensor %s" % name)
        out[name] = np.frombuffer(raw, dtype="<f4").reshape(e["shape"]).copy()
    return out


def load_mccnn_weights(prefix, num_conv_layers=5, verify_crc=True):
    """(weights, biases) lists in layer order, HWIO float32, from ``convN/{weights,biases}``."""
    tensors = load_checkpoint(prefix, verify_crc=verify_crc)
    ws, bs = [], []
    for i in range(1, num_conv_layers + 1):
        ws.append(tensors["conv%d/weights" % i])
        bs.append(tensors["conv%d/biases" % i])
    return ws, bs
