"""Interactions::to_compressed on the device (data_prep.cu) against the host path, which is pinned on the
reference's fixtures in test_oracle_pins.py / test_abi_and_host.py.  The reference sorts STABLY on (user, timestamp)
(data.rs:240): inputs are built with many tied (user, timestamp) pairs so that the tie order is exercised, as in
ML-100K (50,561 tied pairs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _both(pkg, u, i, t, nu, ni):
    host = pkg.CompressedInteractions._from_triplets(u, i, t, nu, ni)
    dev = pkg.CompressedInteractions._from_triplets(u, i, t, nu, ni, device=True)
    return host.arrays(), dev.arrays(), dev


@pytest.mark.parametrize("nnz,nu,ni,tmax", [(0, 5, 7, 3), (1, 3, 4, 9), (100, 20, 20, 50), (50_000, 700, 1683, 40), (300_000, 944, 1683, 2 ** 40)])
def test_device_csr_equals_host_csr(pkg, nnz, nu, ni, tmax):
    rng = np.random.default_rng(nnz + 1)
    u = rng.integers(0, nu, size=nnz).astype(np.uint64)
    i = rng.integers(0, ni, size=nnz).astype(np.uint64)
    t = rng.integers(0, tmax, size=nnz).astype(np.uint64)
    (hp, hi, ht), (dp, di, dt), dev = _both(pkg, u, i, t, nu, ni)
    assert np.array_equal(hp, dp) and np.array_equal(hi, di) and np.array_equal(ht, dt)
    assert len(dev) == nnz and dev.num_users() == nu and dev.num_items() == ni


def test_device_csr_on_ml100k_and_fit(pkg, ml100k):
    """ML-100K's shuffled triples: user pointers and timestamps of the golden CSR (made from the reference's data.csv)
    are reproduced, host and device agree on the order of the 50,561 tied (user, timestamp) pairs, and a model trained
    on the device-built CSR equals one trained on the host-built CSR bit for bit"""
    up = ml100k["user_ptr"].astype(np.int64)
    users = np.repeat(np.arange(944), np.diff(up)).astype(np.uint64)
    items, ts = ml100k["item_ids"].astype(np.uint64), ml100k["timestamps"].astype(np.uint64)
    perm = np.argsort(np.random.default_rng(0).integers(0, 1000, size=len(users)), kind="stable")
    (hp, hi, ht), (dp, di, dt), dev = _both(pkg, users[perm], items[perm], ts[perm], 944, 1683)
    assert np.array_equal(dp, ml100k["user_ptr"].astype(np.uint64)) and np.array_equal(dt, ts)
    assert np.array_equal(hp, dp) and np.array_equal(hi, di) and np.array_equal(ht, dt)
    host = pkg.CompressedInteractions._from_triplets(users[perm], items[perm], ts[perm], 944, 1683)
    outs = []
    for data in (host, dev):
        m = (pkg.ewma.Hyperparameters(1683, 32).embedding_dim(32).learning_rate(0.05).loss(pkg.Loss.BPR)
             .optimizer(pkg.Optimizer.Adagrad).num_epochs(1).num_threads(1).from_seed(bytes(range(16))).build())
        m.fit(data)
        outs.append(m.get_parameter("item_embeddings"))
    assert np.array_equal(outs[0], outs[1])


def test_device_csr_rejects_out_of_range_ids(pkg):
    u = np.array([0, 5], dtype=np.uint64); i = np.array([1, 2], dtype=np.uint64); t = np.array([0, 1], dtype=np.uint64)
    with pytest.raises(pkg.SbrError):
        pkg.CompressedInteractions._from_triplets(u, i, t, 3, 10, device=True)
