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


def _ragged_csr(rng, num_users, num_items, max_len, empty_frac=0.2):
    lens = rng.integers(0, max_len, size=num_users)
    lens[rng.random(num_users) < empty_frac] = 0
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ids = rng.integers(0, num_items, size=int(ptr[-1])).astype(np.uint64)
    return ptr, ids


@pytest.mark.parametrize("num_users,max_len,T,threads", [(1, 40, 8, 1), (50, 12, 3, 2), (3000, 200, 32, 16), (20000, 70, 7, 128),
                                                         (400, 5000, 128, 8), (1000, 30, 2, 1), (300000, 40, 32, 0)])
def test_device_chunker_equals_host_schedule(pkg, num_users, max_len, T, threads):
    """sequence_model.rs:76-84: fit() builds the sub-sequences on the device (data_prep.cu device_schedule) and shuffles their
    indices on the host; sbr_host_schedule (pinned on the oracle and on the reference's chunk golden vector in the CPU suite)
    does all of it on the host.  Same chunks in the same (user) order -- first chunk the short one, len <= 2 dropped, empty
    users skipped -- and the same master-shuffled, partition-major order, remainder dropped."""
    rng = np.random.default_rng(num_users + T)
    ptr, ids = _ragged_csr(rng, num_users, 500, max_len)
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=500)
    model = (pkg.ewma.Hyperparameters(500, T).embedding_dim(16).num_epochs(1).num_threads(threads).from_seed(bytes(range(16))).build())
    state = model.rng_state
    if T <= 2:   # every chunk has <= 2 items: nothing survives the filter (sequence_model.rs:86-88)
        with pytest.raises(pkg.NoInteractions):
            model.fit_plan(data)
        return
    hs, hl, ho, _ = data.host_schedule(T, state)
    plan = model.fit_plan(data)
    ds, dl, do = plan.read_schedule()
    assert np.array_equal(hs, ds) and np.array_equal(hl, dl)
    P = plan.stats()["partitions"] if threads == 0 else threads
    n = len(hs) // P
    assert len(do) == P * n and np.array_equal(ho[:P * n], do)


def test_fit_from_page_locked_host_buffers_equals_pageable(pkg):
    """The id stream of a borrowed CSR is DMA'ed as raw 64-bit words and narrowed on the device when the caller's buffer is
    page-locked, narrowed by host threads into a pinned staging buffer otherwise: same model either way, and an
    out-of-range id is reported on both paths."""
    import torch
    rng = np.random.default_rng(3)
    S, L, N = 4096, 32, 1683
    ptr = (np.arange(S + 1) * L).astype(np.uint64)
    ids = rng.integers(1, N, size=S * L).astype(np.uint64)
    pinned = torch.empty(S * L, dtype=torch.int64).pin_memory()
    ids_pinned = pinned.numpy().view(np.uint64)
    ids_pinned[:] = ids
    outs = []
    for arr in (ids, ids_pinned):
        m = (pkg.ewma.Hyperparameters(N, L).embedding_dim(32).learning_rate(0.05).loss(pkg.Loss.BPR).optimizer(pkg.Optimizer.Adagrad)
             .num_epochs(1).num_threads(1).from_seed(bytes(range(16))).build())   # one partition: deterministic
        m.fit(pkg.CompressedInteractions.from_csr(ptr, arr, None, num_items=N, borrow=True))
        outs.append((m.get_parameter("item_embeddings"), m.last_fit_stats()["h2d_bytes"]))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert outs[1][1] - outs[0][1] == S * L * 4     # raw words on the wire instead of narrowed ones
    ids_pinned[77] = N + 5
    m = pkg.ewma.Hyperparameters(N, L).embedding_dim(32).num_threads(4).from_seed(bytes(range(16))).build()
    with pytest.raises(pkg.SbrError):
        m.fit(pkg.CompressedInteractions.from_csr(ptr, ids_pinned, None, num_items=N, borrow=True))
