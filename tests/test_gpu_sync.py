"""Parallelism::Synchronous (mod.rs:39-40): round-synchronous schedule with the explicit row exchange
(sync_engine.cu).  Single GPU: must reproduce the oracle's barrier mode -- every thread's gradients from the round-start
parameters, sparse entries applied un-merged, one dense step on the round-summed gradient."""
import numpy as np
import pytest

from helpers import make_pair, max_abs_diff, state_names

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,loss,P", [(32, "bpr", 4), (128, "hinge", 8), (16, "bpr", 3)])
def test_synchronous_rounds_match_oracle(pkg, oracle, D, loss, P):
    rng = np.random.default_rng(21)
    N, T, U = 2_000_000, 10, 40      # huge catalogue: two partitions of a round practically never name the same row
    lens = rng.integers(3, 25, size=U)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ids = rng.integers(1, N, size=int(ptr[-1])).astype(np.uint64)
    gm, om = make_pair(pkg, oracle, "ewma", N, T, D, loss=loss, optimizer="adagrad", lr=0.05, l2=1e-3, epochs=2, threads=P,
                       parallelism="synchronous")
    r = np.random.default_rng(3)
    touched = np.unique(ids)
    e = gm.get_parameter("item_embeddings").reshape(N, D)
    e[touched] = (r.standard_normal((len(touched), D)) * 0.3).astype(np.float32)   # meaningful gradients on the rows in use
    gm.set_parameter("item_embeddings", e)
    gm.set_parameter("alpha", (r.standard_normal(D) * 0.5).astype(np.float32))
    for n in om.param_names():
        om.param(n)[:] = gm.get_parameter(n)
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N)
    gl = gm.fit(data)
    rc, ol = om.fit(ptr, ids)
    assert rc == 0
    st = gm.last_fit_stats()
    assert st["partitions"] == P and st["kernel_launches"] > 2 * (st["steps"] // P)   # one set of kernels per round
    diffs = max_abs_diff(gm, om, state_names(om, "adagrad"))
    assert max(diffs.values()) <= 2e-4, diffs
    assert abs(gl - ol) <= 1e-4 * max(1.0, abs(ol))
    assert gm.num_updates == om.num_updates and gm.rng_state == om.rng_state


def test_synchronous_is_deterministic_and_learns(pkg):
    rng = np.random.default_rng(5)
    N, T, D = 1683, 32, 32
    ptr = (np.arange(4097) * 32).astype(np.uint64)
    ids = rng.integers(1, N, size=4096 * 32).astype(np.uint64)
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N)
    losses = []
    for _ in range(2):
        m = (pkg.ewma.Hyperparameters(N, T).embedding_dim(D).learning_rate(0.05).loss(pkg.Loss.BPR).optimizer(pkg.Optimizer.Adagrad)
             .parallelism(pkg.Parallelism.Synchronous).num_threads(64).num_epochs(1).from_seed(bytes(range(16))).build())
        losses.append([m.fit(data) / 64 for _ in range(3)])
        assert np.all(np.isfinite(m.get_parameter("item_embeddings")))
    assert losses[0][-1] < losses[0][0]
    assert abs(losses[0][0] - losses[1][0]) < 1e-3   # same schedule; only colliding rows inside a round may race


@pytest.mark.parametrize("D,P", [(32, 4), (64, 6)])
def test_synchronous_ewma_warp_matches_oracle(pkg, oracle, D, P):
    """WARP under Parallelism::Synchronous on one GPU: the requested negative is candidate 0, further candidates are scored
    against the round-start table (sequence_model.rs:47-68), the accepted one replaces the request before the apply stage."""
    rng = np.random.default_rng(33)
    N, T, U = 1_000_000, 10, 40
    lens = rng.integers(3, 25, size=U)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ids = rng.integers(1, N, size=int(ptr[-1])).astype(np.uint64)
    gm, om = make_pair(pkg, oracle, "ewma", N, T, D, loss="warp", optimizer="adagrad", lr=0.05, l2=1e-3, epochs=2, threads=P,
                       parallelism="synchronous")
    r = np.random.default_rng(3)
    touched = np.unique(ids)
    e = gm.get_parameter("item_embeddings").reshape(N, D)
    e[touched] = (r.standard_normal((len(touched), D)) * 0.3).astype(np.float32)
    gm.set_parameter("item_embeddings", e)
    gm.set_parameter("item_biases", (r.standard_normal(N) * 0.8).astype(np.float32))   # spread biases: candidates do get rejected
    gm.set_parameter("alpha", (r.standard_normal(D) * 0.5).astype(np.float32))
    for n in om.param_names():
        om.param(n)[:] = gm.get_parameter(n)
    b0 = gm.get_parameter("item_biases").copy()
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N)
    gl = gm.fit(data)
    rc, ol = om.fit(ptr, ids)
    assert rc == 0
    assert "round-synchronous" in gm.last_fit_stats()["kernel"]
    diffs = max_abs_diff(gm, om, state_names(om, "adagrad"))
    assert max(diffs.values()) <= 2e-4, diffs
    assert abs(gl - ol) <= 1e-4 * max(1.0, abs(ol))
    # the resampling did happen: more negative rows were visited than one per timestep could explain without rejections
    moved = np.flatnonzero(om.param("item_biases") != b0)
    assert len(moved) > len(touched)


def test_synchronous_never_falls_back_to_hogwild_silently(pkg):
    """A configuration the round-synchronous engines cannot run (WARP on a row-sharded table) is refused, not run asynchronously."""
    rng = np.random.default_rng(1)
    ptr = (np.arange(65) * 8).astype(np.uint64)
    ids = rng.integers(1, 100, size=64 * 8).astype(np.uint64)
    m = (pkg.ewma.Hyperparameters(100, 8).embedding_dim(32).loss(pkg.Loss.WARP).parallelism(pkg.Parallelism.Synchronous).num_threads(4)
         .from_seed(bytes(range(16))).virtual_shards(2).build())
    with pytest.raises(pkg.SbrError):
        m.fit(pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=100))


@pytest.mark.parametrize("loss,P", [("bpr", 8), ("warp", 16)])
def test_synchronous_rounds_with_colliding_rows_match_oracle(pkg, oracle, loss, P):
    """A catalogue so small that several partitions of a round name the same rows (as on ML-100K): the entries of a row are
    applied un-merged in the reference order -- thread-major, then t descending, E[neg], E[out], E[in] -- so the result is still
    the oracle's barrier mode element for element (optimizer state started at 1: see tests/test_gpu_lstm_batch.py)."""
    rng = np.random.default_rng(17)
    N, T, D, U = 120, 10, 32, 64
    lens = rng.integers(3, 25, size=U)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ids = rng.integers(1, N, size=int(ptr[-1])).astype(np.uint64)
    gm, om = make_pair(pkg, oracle, "ewma", N, T, D, loss=loss, optimizer="adagrad", lr=0.05, l2=1e-3, epochs=2, threads=P,
                       parallelism="synchronous", scale=0.3)
    for n in om.param_names():
        gm.set_parameter(n + ".s1", np.ones(len(gm.get_parameter(n)), dtype=np.float32))
    for n in state_names(om, "adagrad"):
        om.param(n)[:] = gm.get_parameter(n)
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N)
    gl = gm.fit(data)
    rc, ol = om.fit(ptr, ids)
    assert rc == 0
    diffs = max_abs_diff(gm, om, state_names(om, "adagrad"))
    assert max(diffs.values()) <= 3e-4, diffs
    assert abs(gl - ol) <= 1e-4 * max(1.0, abs(ol))
