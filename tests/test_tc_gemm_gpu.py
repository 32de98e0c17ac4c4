"""GPU suite: the bf16 tcgen05 GEMM of the training step (az_tc_gemm.cu) through its C-ABI test entry against numpy on the same
bf16-rounded operands (fp32 products are exact; only the accumulation order differs)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from alphazero_risk_b200 import api as a
    if a.lib().az_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu suite must run on a B200")
    a.lib().az_tc_gemm_test.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_float)]
    return a


def bf16_round(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7fff + ((u >> 16) & 1)) & 0xffff0000          # round to nearest even
    return u.astype(np.uint32).view(np.float32)


def gemm(api, a, b, splits, reps=1):
    M, K = a.shape
    c = np.empty((M, 256), np.float32)
    ms = C.c_float(0)
    api.check(api.lib().az_tc_gemm_test(a.ctypes.data, b.ctypes.data, M, K, splits, reps, c.ctypes.data, C.byref(ms)))
    return c, float(ms.value)


@pytest.mark.parametrize("M,K,splits", [(128, 64, 1), (128, 256, 1), (300, 200, 1), (256, 1024, 4), (1000, 2304, 1), (117, 4000, 8), (2304, 3000, 16)])
def test_gemm_matches_numpy(api, M, K, splits):
    rng = np.random.default_rng(M * 7 + K)
    a = bf16_round(rng.standard_normal((M, K)).astype(np.float32))
    b = bf16_round(rng.standard_normal((256, K)).astype(np.float32))
    c, _ = gemm(api, a, b, splits)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    err = np.abs(c - ref).max() / np.abs(ref).max()
    assert err < 3e-5, (M, K, splits, err)        # fp32 accumulation over up to 4000 terms


def test_gemm_structured_operands_catch_layout_mistakes(api):
    """A = one-hot rows, B = distinct integers: every output element identifies the (m, n, k) triple that produced it"""
    M, K = 256, 192
    a = np.zeros((M, K), np.float32)
    a[np.arange(M), (np.arange(M) * 5) % K] = 1.0
    b = (np.arange(256)[:, None] + 2.0 * (np.arange(K)[None, :] % 64)).astype(np.float32)        # exactly representable in bf16 (< 512, even steps)
    b = bf16_round(b)
    c, _ = gemm(api, a, b, 2)
    assert (c == a @ b.T).all()


def test_gemm_throughput_at_the_training_shapes(api, capsys):
    rng = np.random.default_rng(1)
    out = []
    for name, M, K, splits in (("forward / data gradient, batch 512", 21504, 2304, 1), ("weight gradient, batch 512", 2304, 21504, 8)):
        a = bf16_round(rng.standard_normal((M, K)).astype(np.float32) * 0.1)
        b = bf16_round(rng.standard_normal((256, K)).astype(np.float32) * 0.1)
        c, ms = gemm(api, a, b, splits, reps=5)
        tf = 2.0 * M * 256 * K / (ms * 1e-3) / 1e12
        out.append("%s: M %d K %d splits %d: %.1f us, %.0f TFLOP/s" % (name, M, K, splits, ms * 1e3, tf))
        idx = rng.integers(0, M, 64)
        ref = a[idx].astype(np.float64) @ b.astype(np.float64).T
        assert np.abs(c[idx] - ref).max() / np.abs(ref).max() < 1e-4
        assert tf > 100.0, out[-1]
    with capsys.disabled():
        print("\n" + "\n".join(out))
