"""TEST INFRASTRUCTURE — pure-Python restatement of TensorFlow's V2 checkpoint bundle format, the checker for
alphazero_risk_b200/csrc/az_ckpt.cpp (AlphaZeroNN::saveCheckpoint / loadCheckpoint, reference
src/risk_game/player/alpha_zero/neural_network/alphazero_nn.cpp:189-214, exchange these files through the graph's Saver).

Published format (tensorflow/core/util/tensor_bundle/tensor_bundle.cc, tensorflow/core/lib/io/{table_builder,block_builder,format}.cc,
tensorflow/core/protobuf/tensor_bundle.proto, tensorflow/core/lib/hash/crc32c.h):
  <prefix>.index = LevelDB-style table: data blocks, meta-index block, index block, 48-byte footer (two BlockHandles padded to 40
  bytes + magic 0xdb4775248b80fb57); block = prefix-compressed entries + restart array + count, then 1 type byte + masked CRC32C;
  key "" -> BundleHeaderProto, key name -> BundleEntryProto{dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32)}.
  <prefix>.data-00000-of-00001 = raw little-endian tensor bytes.

PARITY UNPINNED: TensorFlow is absent (un-vendored, un-pinned dependency of the reference) and the reference ships no checkpoint.
What IS pinned: CRC32C known answers (RFC 3720 B.4), the tensor inventory of the shipped GraphDef (tests/golden/ckpt_tensors_V2_5.json,
generated from python/model/model_txt_V2_5.pb by tests/golden/gen_ckpt_tensors.py).  Only tests/ may import this module.
"""
import os
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
MASK_DELTA = 0xa282ead8

_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82f63b78 if _c & 1 else _c >> 1
    _TABLE.append(_c)


def _c_crc():
    """the bit-serial C restatement in oracle/risk_oracle.c (fast path for tensor data); None when the oracle library is not built"""
    global _CLIB
    if _CLIB is None:
        try:
            import ctypes
            lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "librisk_oracle.so"))
            lib.ro_crc32c.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint32]
            lib.ro_crc32c.restype = ctypes.c_uint32
            _CLIB = lib
        except (OSError, AttributeError):
            _CLIB = False
    return _CLIB or None


_CLIB = None


def crc32c(data, crc=0):
    data = bytes(data)
    if len(data) > 4096 and _c_crc() is not None:
        return int(_c_crc().ro_crc32c(data, len(data), crc))
    c = crc ^ 0xffffffff
    for b in data:
        c = _TABLE[(c ^ b) & 0xff] ^ (c >> 8)
    return c ^ 0xffffffff


def mask(c):
    return (((c >> 15) | (c << 17)) + MASK_DELTA) & 0xffffffff


def unmask(m):
    r = (m - MASK_DELTA) & 0xffffffff
    return ((r >> 17) | (r << 15)) & 0xffffffff


def _varint(v):
    out = bytearray()
    while v >= 128:
        out.append((v & 127) | 128)
        v >>= 7
    out.append(v)
    return bytes(out)


def _get_varint(buf, pos):
    v, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 127) << shift
        if not b & 128:
            return v, pos
        shift += 7


def _fields(buf):
    """protobuf wire format -> [(field, wire, value)]"""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        f, w = tag >> 3, tag & 7
        if w == 0:
            v, pos = _get_varint(buf, pos)
        elif w == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif w == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]; pos += n
        elif w == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError("wire type %d" % w)
        out.append((f, w, v))
    return out


def _read_block(data, off, size):
    body = data[off:off + size]
    assert data[off + size] == 0, "compressed block"
    (stored,) = struct.unpack("<I", data[off + size + 1:off + size + 5])
    assert unmask(stored) == crc32c(data[off:off + size + 1]), "block checksum"
    (n_restarts,) = struct.unpack("<I", body[-4:])
    end = size - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(body, pos)
        unshared, pos = _get_varint(body, pos)
        vlen, pos = _get_varint(body, pos)
        key = key[:shared] + body[pos:pos + unshared]; pos += unshared
        out.append((key, body[pos:pos + vlen])); pos += vlen
    return out


def read_bundle(prefix):
    """-> {name: float32 array} (verifies every checksum)"""
    index = open(prefix + ".index", "rb").read()
    foot = index[-48:]
    assert struct.unpack("<Q", foot[40:])[0] == MAGIC
    pos = 0
    _, pos = _get_varint(foot, pos); _, pos = _get_varint(foot, pos)
    io, pos = _get_varint(foot, pos); isz, pos = _get_varint(foot, pos)
    kv = []
    for _, handle in _read_block(index, io, isz):
        bo, p = _get_varint(handle, 0); bs, p = _get_varint(handle, p)
        kv += _read_block(index, bo, bs)
    assert kv[0][0] == b""
    header = {f: v for f, w, v in _fields(kv[0][1])}
    assert header.get(1, 0) == 1 and header.get(2, 0) == 0, "one shard, little endian"
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    out = {}
    names = [k for k, _ in kv[1:]]
    assert names == sorted(names)
    for key, val in kv[1:]:
        e = {f: v for f, w, v in _fields(val)}
        assert e[1] == 1, "DT_FLOAT"
        shape = []
        for f, w, v in _fields(e.get(2, b"")):
            if f == 2:
                d = {ff: vv for ff, ww, vv in _fields(v)}
                shape.append(d.get(1, 0))
        off, size = e.get(4, 0), e[5]
        raw = data[off:off + size]
        assert unmask(struct.unpack("<I", e[6])[0]) == crc32c(raw), key
        out[key.decode()] = np.frombuffer(raw, "<f4").reshape(shape).copy()
    return out


def write_bundle(prefix, tensors, block_size=262144, restart_interval=16, hostile_entries=None):
    """{name: array} -> the two files, entries and data in ascending name order.  hostile_entries = {name: (offset, size)} writes
    those (out-of-range) values into the entry instead of the true ones: for the reader's bounds-check tests"""
    blob = bytearray()
    entries = [(b"", b"\x08\x01" + b"\x1a\x02\x08\x01")]          # num_shards = 1, version { producer = 1 }
    for name in sorted(tensors):
        a = np.asarray(tensors[name], dtype="<f4", order="C")        # (ascontiguousarray would turn a scalar into shape (1,))
        raw = a.tobytes()
        shape = b"".join(b"\x12" + _varint(len(d)) + d for d in (b"\x08" + _varint(s) for s in a.shape))
        e = b"\x08\x01" + b"\x12" + _varint(len(shape)) + shape
        off, size = (hostile_entries or {}).get(name, (len(blob), len(raw)))
        if off:
            e += b"\x20" + _varint(off)
        e += b"\x28" + _varint(size) + b"\x35" + struct.pack("<I", mask(crc32c(raw)))
        entries.append((name.encode(), e))
        blob += raw

    def build_block(items):
        buf, restarts, last, counter = bytearray(), [0], b"", 0
        for k, v in items:
            shared = 0
            if counter < restart_interval:
                while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                    shared += 1
            else:
                restarts.append(len(buf)); counter = 0
            buf += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
            last = k; counter += 1
        for r in restarts:
            buf += struct.pack("<I", r)
        return bytes(buf + struct.pack("<I", len(restarts)))

    table, index_items, cur, cur_size = bytearray(), [], [], 0

    def emit(block):
        handle = _varint(len(table)) + _varint(len(block))
        table.extend(block + b"\x00" + struct.pack("<I", mask(crc32c(block + b"\x00"))))
        return handle

    for k, v in entries:
        cur.append((k, v)); cur_size += len(k) + len(v) + 3
        if cur_size >= block_size:
            index_items.append((cur[-1][0], emit(build_block(cur)))); cur, cur_size = [], 0
    if cur:
        index_items.append((cur[-1][0], emit(build_block(cur))))
    foot = emit(build_block([])) + emit(build_block(index_items))
    foot += b"\x00" * (40 - len(foot)) + struct.pack("<Q", MAGIC)
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(blob))
    open(prefix + ".index", "wb").write(bytes(table) + foot)
