"""Round-synchronous batched LSTM engine (lstm_batch.cuh: tcgen05 GEMMs over 128 x 128 bf16 tiles) against the CPU oracle.

Parallelism::Synchronous (mod.rs:39-40; the reference default, lstm.rs:66) with num_threads > 1: every thread's gradients
come from the round-start parameters, the sparse entries are applied un-merged in thread order, the dense LSTM weights
take ONE step on the gradient summed over the round -- the oracle's barrier mode (oracle/sbr_oracle.c run_partition).
The engine's products run on bf16 operands with fp32 accumulation (h_{t-1}, x_t, W and the gate deltas are rounded to
bf16: 2^-9 relative per element); scores, losses, the cell recurrence and every optimizer step are fp32.  Stated
tolerance: parameters within 1.5e-3 of the oracle after a whole fit of several rounds at lr 0.05 (updates are O(1e-2) per
round), second-moment state within 2 % + 5e-3, Adam's first moment within 2 % + 2e-3.  The optimizer's second-moment state starts at 1 in these tests: at 0 the
first Adagrad / Adam step is lr * sign(g) whatever |g| is, which turns a 1-ulp difference into an O(lr) one.
"""
import numpy as np
import pytest

from helpers import make_pair, state_names

pytestmark = pytest.mark.gpu

CASES = [  # D, variant, loss, optimizer, partitions
    (16, "coupled", "bpr", "adam", 4),        # the reference's default Hyperparameters (lstm.rs:56-71) but for num_threads
    (32, "normal", "bpr", "adagrad", 8),
    (32, "normal", "hinge", "adagrad", 3),
    (64, "normal", "hinge", "adam", 8),       # BASELINE config C3's model
    (64, "coupled", "bpr", "adagrad", 5),
    (128, "normal", "bpr", "adagrad", 4),
    (256, "normal", "bpr", "adagrad", 3),     # BASELINE config C5's width
]


def _prepare(pkg, oracle, D, variant, loss, optimizer, P, N, T, ptr, ids, epochs=2, lr=0.05):
    gm, om = make_pair(pkg, oracle, "lstm", N, T, D, loss=loss, optimizer=optimizer, variant=variant, lr=lr, l2=1e-3, epochs=epochs,
                       threads=P, parallelism="synchronous")
    r = np.random.default_rng(3)
    touched = np.unique(ids)
    e = gm.get_parameter("item_embeddings").reshape(N, D)
    e[touched] = (r.standard_normal((len(touched), D)) * 0.3).astype(np.float32)   # meaningful gradients on the rows in use
    gm.set_parameter("item_embeddings", e)
    b = gm.get_parameter("item_biases")
    b[touched] = (r.standard_normal(len(touched)) * 0.3).astype(np.float32)
    gm.set_parameter("item_biases", b)
    slot = ".s2" if optimizer == "adam" else ".s1"
    for n in om.param_names():
        gm.set_parameter(n + slot, np.ones(len(gm.get_parameter(n)), dtype=np.float32))
    for n in state_names(om, optimizer):
        om.param(n)[:] = gm.get_parameter(n)
    return gm, om


@pytest.mark.parametrize("D,variant,loss,optimizer,P", CASES)
def test_synchronous_lstm_rounds_match_oracle(pkg, oracle, D, variant, loss, optimizer, P):
    rng = np.random.default_rng(21 + D)
    N, T, U = 2_000_000, 10, 40      # huge catalogue: two partitions of a round practically never name the same row
    lens = rng.integers(3, 25, size=U)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ids = rng.integers(1, N, size=int(ptr[-1])).astype(np.uint64)
    gm, om = _prepare(pkg, oracle, D, variant, loss, optimizer, P, N, T, ptr, ids)
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N)
    e0 = gm.get_parameter("item_embeddings").reshape(N, D).copy()
    gl = gm.fit(data)
    rc, ol = om.fit(ptr, ids)
    assert rc == 0
    st = gm.last_fit_stats()
    assert st["partitions"] == P and "batched" in st["kernel"]
    assert st["kernel_launches"] > 4 * (T - 1) * (st["steps"] // P)     # per round: (T - 1) x {fwd, score, delta, dz} + the rest
    assert gm.num_updates == om.num_updates and gm.rng_state == om.rng_state
    # rows the oracle visited (sequence items and the negatives of the shared counter-based sampler)
    touched = np.flatnonzero(np.any(om.param("item_embeddings").reshape(N, D) != e0, axis=1))
    assert len(touched) >= len(np.unique(ids))
    for n in state_names(om, optimizer):
        a, b = gm.get_parameter(n), om.param(n)
        assert np.all(np.isfinite(a)), n
        if n.startswith("item_embeddings"):
            a, b = a.reshape(N, D)[touched], b.reshape(N, D)[touched]
        elif n.startswith("item_biases"):
            a, b = a[touched], b[touched]
        if n.endswith(".s1") and optimizer == "adagrad" or n.endswith(".s2"):
            assert np.all(np.abs(a - b) <= 0.02 * np.abs(b - 1.0) + 5e-3), (n, float(np.max(np.abs(a - b))))
        elif n.endswith(".s1"):   # Adam's first moment: a decaying sum of gradients, held to the gradients' relative accuracy
            assert np.all(np.abs(a - b) <= 0.02 * np.abs(b) + 2e-3), (n, float(np.max(np.abs(a - b))))
        else:
            assert float(np.max(np.abs(a - b))) <= 1.5e-3, (n, float(np.max(np.abs(a - b))))
    moved = np.abs(om.param("lstm_weights") - gm.get_parameter("lstm_weights")).max(), np.abs(gm.get_parameter("lstm_weights")).max()
    assert abs(gl - ol) <= 2e-3 * max(1.0, abs(ol)), (gl, ol, moved)
    # untouched rows are bit-identical to the start (the engine only visits recorded entries)
    e1 = gm.get_parameter("item_embeddings").reshape(N, D)
    mask = np.ones(N, dtype=bool); mask[touched] = False
    assert np.array_equal(e1[mask], e0[mask])


@pytest.mark.parametrize("D,P", [(32, 6), (64, 4)])
def test_synchronous_lstm_warp_negatives_and_parameters(pkg, oracle, D, P):
    """WARP under Synchronous: candidates are scored against the round-start table (sequence_model.rs:47-68).  A rejection
    decision whose margin is within bf16 noise of 0 may fall differently than in the oracle; rows visited by such a
    sequence are few, every other row is held to the stated tolerance."""
    rng = np.random.default_rng(5 + D)
    N, T, U = 500_000, 8, 36
    lens = rng.integers(3, 17, size=U)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ids = rng.integers(1, N, size=int(ptr[-1])).astype(np.uint64)
    gm, om = _prepare(pkg, oracle, D, "normal", "warp", "adagrad", P, N, T, ptr, ids, epochs=1)
    b = gm.get_parameter("item_biases")     # spread biases: candidates do get rejected
    b[:] = (np.random.default_rng(1).standard_normal(N) * 0.8).astype(np.float32)
    gm.set_parameter("item_biases", b); om.param("item_biases")[:] = b
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N)
    gm.fit(data)
    rc, _ = om.fit(ptr, ids)
    assert rc == 0
    ge, oe = gm.get_parameter("item_embeddings").reshape(N, D), om.param("item_embeddings").reshape(N, D)
    e0 = None
    changed_g = np.flatnonzero(np.abs(ge).max(axis=1) > 0)   # every row has a nonzero init; compare the rows either side moved
    diff = np.abs(ge - oe).max(axis=1)
    visited = np.flatnonzero(diff > 0)
    bad = np.flatnonzero(diff > 1.5e-3)
    assert len(visited) > 0.5 * len(np.unique(ids))
    assert len(bad) <= 0.05 * len(visited) + 2, (len(bad), len(visited))
    assert np.abs(gm.get_parameter("lstm_weights") - om.param("lstm_weights")).max() <= 4e-3


def test_wide_lstm_many_partitions_runs_on_the_batched_engine_and_learns(pkg, oracle):
    """Asynchronous fit of a wide LSTM (embedding_dim 64) with >= 128 partitions: rounds of the batched tensor-core engine
    (Hogwild at round granularity); per-epoch losses fall and track the exact fp32 FFMA kernel run with the same
    partition count (no worse than it)."""
    rng = np.random.default_rng(7)
    N, T, D = 5000, 16, 64
    S = 4096
    ptr = (np.arange(S + 1) * T).astype(np.uint64)
    ids = rng.integers(1, N, size=S * T).astype(np.uint64)
    losses = {}
    for kern in ("exact", "batched"):
        gm, _ = make_pair(pkg, oracle, "lstm", N, T, D, loss="hinge", optimizer="adagrad", lr=0.05, l2=1e-4, epochs=1, threads=256,
                          exact=(kern == "exact"))
        data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N)
        losses[kern] = [gm.fit(data) / 256 for _ in range(5)]
        assert ("batched" in gm.last_fit_stats()["kernel"]) == (kern == "batched")
        for n in ("item_embeddings", "item_biases", "lstm_weights", "lstm_biases"):
            assert np.all(np.isfinite(gm.get_parameter(n))), (kern, n)
    # same data, same partition count, different schedules: Hogwild steps the dense weights once per sub-sequence from
    # whatever the other 255 partitions left, the rounds step them once per 256 sub-sequences on the summed gradient --
    # the curves agree at the start and neither may lag far behind the other
    a, b = np.array(losses["exact"]), np.array(losses["batched"])
    assert b[-1] < b[0] - 0.02 and a[-1] < a[0] - 0.02, (a, b)
    assert abs(a[0] - b[0]) < 0.06 and b[-1] < a[-1] + 0.05, (a, b)
