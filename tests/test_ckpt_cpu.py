"""CPU suite: the TensorFlow V2 checkpoint bundle reader/writer of libaz_b200.so (az_ckpt_*, host only) against the pure-Python
restatement of the format (oracle/ckpt_oracle.py) in both directions, CRC32C known answers, corruption detection, and the tensor
inventory the reference graph's Saver writes (tests/golden/ckpt_tensors_V2_5.json, from the shipped GraphDef)."""
import json
import os

import numpy as np
import pytest

from oracle import ckpt_oracle as co


@pytest.fixture(scope="module")
def api():
    from alphazero_risk_b200 import api as a
    a.lib()
    return a


def test_crc32c_known_answers(api):
    # RFC 3720 appendix B.4
    L = api.lib()
    for data, want in ((b"\x00" * 32, 0x8a9136aa), (b"\xff" * 32, 0x62a8ab43), (bytes(range(32)), 0x46dd794e),
                       (bytes(range(31, -1, -1)), 0x113fdb5c), (b"123456789", 0xe3069283)):
        a = np.frombuffer(data, np.uint8)
        assert L.az_crc32c(a.ctypes.data, len(data)) == want
        assert co.crc32c(data) == want
    assert co.unmask(co.mask(0x12345678)) == 0x12345678


def graph_tensors(golden_dir, rng):
    spec = json.load(open(os.path.join(golden_dir, "ckpt_tensors_V2_5.json")))["tensors"]
    return {n: rng.standard_normal(s).astype(np.float32) for n, s in spec}, spec


def test_python_written_bundle_is_read_by_the_library(api, golden_dir, tmp_path):
    tensors, spec = graph_tensors(golden_dir, np.random.default_rng(1))
    prefix = str(tmp_path / "az_train")
    co.write_bundle(prefix, tensors)
    ck = api.Checkpoint(prefix)
    listed = ck.tensors()
    assert [t[0] for t in listed] == sorted(tensors) == [n for n, _ in spec]          # table order = the Saver's sorted names
    for name, dtype, shape, nbytes in listed:
        assert dtype == 1 and shape == tensors[name].shape and nbytes == tensors[name].nbytes
        assert (ck.read(name).view(np.uint32) == tensors[name].view(np.uint32)).all()
    with pytest.raises(KeyError):
        ck.read("no/such/tensor")
    ck.close()


@pytest.mark.parametrize("block_size", [262144, 4096, 300])
def test_library_written_bundle_is_read_by_python_and_multi_block_tables(api, golden_dir, tmp_path, block_size):
    tensors, _ = graph_tensors(golden_dir, np.random.default_rng(2))
    prefix = str(tmp_path / "ck")
    api.Checkpoint.write(prefix, tensors)
    back = co.read_bundle(prefix)
    assert sorted(back) == sorted(tensors)
    for n in tensors:
        assert back[n].shape == tensors[n].shape and (back[n].view(np.uint32) == tensors[n].view(np.uint32)).all()
    # a table TensorFlow could have written with smaller blocks (several data blocks, longer index): same contents through the library
    p2 = str(tmp_path / "small_blocks")
    co.write_bundle(p2, tensors, block_size=block_size, restart_interval=4 if block_size < 1000 else 16)
    ck = api.Checkpoint(p2)
    assert len(ck.tensors()) == len(tensors)
    for n in ("beta1_power", "conv/kernel", "res4e_branch2b/kernel/optimize_1", "v/kernel"):
        assert (ck.read(n).view(np.uint32) == tensors[n].view(np.uint32)).all()
    ck.close()


def test_both_writers_produce_identical_files(api, golden_dir, tmp_path):
    tensors, _ = graph_tensors(golden_dir, np.random.default_rng(3))
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    api.Checkpoint.write(a, tensors)
    co.write_bundle(b, tensors)
    for ext in (".index", ".data-00000-of-00001"):
        assert open(a + ext, "rb").read() == open(b + ext, "rb").read(), ext


def test_corruption_is_detected(api, tmp_path):
    rng = np.random.default_rng(4)
    tensors = {"a/kernel": rng.standard_normal((3, 5)).astype(np.float32), "b": np.float32(2.5).reshape(())}
    prefix = str(tmp_path / "c")
    api.Checkpoint.write(prefix, tensors)
    assert float(api.Checkpoint(prefix).read("b")) == 2.5
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[7] ^= 0x10
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    ck = api.Checkpoint(prefix)
    with pytest.raises(api.AzError, match="checksum"):
        ck.read("a/kernel")
    ck.close()
    index = bytearray(open(prefix + ".index", "rb").read())
    index[5] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(index))
    with pytest.raises(api.AzError, match="checksum|corrupt|bad"):
        api.Checkpoint(prefix)
    with pytest.raises(api.AzError):
        api.Checkpoint(str(tmp_path / "missing"))
    with pytest.raises(api.AzError, match="duplicate"):
        L = api.lib()
        import ctypes as C
        names = (C.c_char_p * 2)(b"x", b"x"); ranks = (C.c_int * 2)(0, 0)
        one = np.zeros(1, np.float32)
        shp = (C.POINTER(C.c_int64) * 2)(); dat = (C.c_void_p * 2)(one.ctypes.data, one.ctypes.data)
        api.check(L.az_ckpt_write(str(tmp_path / "d").encode(), 2, names, ranks, shp, dat))


def test_hostile_offsets_do_not_wrap_the_bounds_checks(api, tmp_path):
    """values read from the file are never added before they are checked: an entry offset / block handle near 2^64 (or a negative
    varint offset) must be rejected, not wrapped past the test into an out-of-bounds read"""
    import struct
    from oracle import ckpt_oracle as co
    t = {"w": np.arange(6, dtype=np.float32)}
    for i, (off, size) in enumerate([(2 ** 64 - 10, 24), (2 ** 63 + 5, 24), (2 ** 64 - 24, 24), (1 << 40, 24)]):
        prefix = str(tmp_path / ("h%d" % i))
        co.write_bundle(prefix, t, hostile_entries={"w": (off, size)})
        ck = api.Checkpoint(prefix)
        with pytest.raises(api.AzError, match="outside"):
            ck.read("w")
        ck.close()
    prefix = str(tmp_path / "f")
    co.write_bundle(prefix, t)
    idx = bytearray(open(prefix + ".index", "rb").read())
    foot = co._varint(0) + co._varint(8) + co._varint(2 ** 64 - 10) + co._varint(20)       # index handle wraps off + size + 5
    idx[-48:] = foot + b"\x00" * (40 - len(foot)) + bytes(idx[-8:])
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(api.AzError, match="outside|corrupt|bad|checksum"):
        api.Checkpoint(prefix)
