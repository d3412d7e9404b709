"""tcgen05 / TMEM tile primitives (sbr-rs_b200/csrc/tc_tile.cuh) against numpy: the three GEMM shapes of the
tensor-core LSTM kernel, including the dual K-major / MN-major view of one shared-memory tile."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tf32(x):
    u = x.view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def _bf16(x):
    u = x.view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(np.float32)


def test_tile_gemms(pkg):
    import os
    L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libtc_primitives_hook.so"))
    f32p = C.POINTER(C.c_float)
    L.sbrdbg_tc_gemm.argtypes = [C.c_int, f32p, f32p, f32p, f32p]
    rng = np.random.default_rng(0)
    Z = rng.standard_normal((128, 80)).astype(np.float32)
    Dl = rng.standard_normal((128, 128)).astype(np.float32)
    W = rng.standard_normal((64, 128)).astype(np.float32)
    Zt, Wt = _tf32(Z.copy()), _tf32(W.copy())
    Zb, Db, Wb = _bf16(Z.copy()), _bf16(Dl.copy()), _bf16(W.copy())
    refs = {1: Zt[:, :64].astype(np.float64) @ Wt.astype(np.float64),        # gates, tf32
            2: Db.astype(np.float64) @ Wb.astype(np.float64).T,               # dz, bf16 K-major
            3: Db.astype(np.float64).T @ Zb.astype(np.float64)}               # dW^T, bf16 MN-major views
    for mode, ref in refs.items():
        out = np.zeros(ref.shape, dtype=np.float32)
        rc = L.sbrdbg_tc_gemm(mode, Z.ctypes.data_as(f32p), Dl.ctypes.data_as(f32p), W.ctypes.data_as(f32p),
                              out.ctypes.data_as(f32p))
        assert rc == 0
        assert np.abs(out - ref).max() < 1e-4, (mode, np.abs(out - ref).max())  # exact products, fp32 accumulation
