"""Row-sharded item table: the sharded addressing (shard = id % G, row = id / G) must be invisible.
Single-GPU part: G "virtual shards" in one process give bit-identical results to the unsharded model.
The real multi-process / multi-GPU path (CUDA-IPC peer mappings over NVLink) is exercised by tests/dist_worker.py under
torchrun on a multi-GPU box (tests/test_gpu_multi.py)."""
import numpy as np
import pytest

from helpers import LOSSES, OPTS, VARIANTS, random_csr

pytestmark = pytest.mark.gpu


def _build(pkg, kind, N, T, D, shards, loss, opt, threads=1, seed=bytes(range(3, 19))):
    H = pkg.lstm.Hyperparameters if kind == "lstm" else pkg.ewma.Hyperparameters
    h = (H(N, T).embedding_dim(D).learning_rate(0.05).l2_penalty(1e-3).loss(LOSSES[loss]).optimizer(OPTS[opt])
         .num_epochs(2).num_threads(threads).parallelism(0).from_seed(seed).virtual_shards(shards))   # Parallelism::Asynchronous
    if kind == "lstm":
        h = h.lstm_variant(VARIANTS["normal"])
    return h.build()


@pytest.mark.parametrize("kind,D,loss,opt", [("ewma", 32, "warp", "adagrad"), ("ewma", 128, "bpr", "adam"),
                                             ("lstm", 32, "warp", "adagrad"), ("lstm", 16, "hinge", "adam")])
@pytest.mark.parametrize("shards", [2, 8])
def test_virtual_shards_bit_identical(pkg, kind, D, loss, opt, shards):
    rng = np.random.default_rng(1)
    N, T = 301, 12  # N not a multiple of the shard count: ragged shards
    ptr, ids = random_csr(rng, 25, N, 1, 40, first_item=0)
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N)
    ref = _build(pkg, kind, N, T, D, 1, loss, opt)
    shd = _build(pkg, kind, N, T, D, shards, loss, opt)
    names = ["item_embeddings", "item_biases"] + (["lstm_weights", "lstm_biases"] if kind == "lstm" else ["alpha"])
    for n in names:  # identical init (counter-based by global row id) ...
        assert np.array_equal(ref.get_parameter(n), shd.get_parameter(n)), n
    probe = rng.integers(0, N, size=500).astype(np.uint64)
    assert np.array_equal(ref.gather_rows(probe).view(np.uint32), shd.gather_rows(probe).view(np.uint32))
    l0, l1 = ref.fit(data), shd.fit(data)
    assert l0 == l1
    for n in names:  # ... and identical training, state included
        for suffix in ("", ".s1") + ((".s2",) if opt == "adam" else ()):
            a, b = ref.get_parameter(n + suffix), shd.get_parameter(n + suffix)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), n + suffix
    assert pkg.mrr_score(ref, data) == pkg.mrr_score(shd, data)
    # set_parameter scatters into the shards
    e = rng.standard_normal(N * D).astype(np.float32)
    shd.set_parameter("item_embeddings", e)
    assert np.array_equal(shd.get_parameter("item_embeddings"), e)
    assert np.array_equal(shd.gather_rows(probe), e.reshape(N, D)[probe.astype(np.int64)])


def test_sharded_tile_kernel_runs(pkg):
    """The tensor-core LSTM kernel also goes through the sharded addressing."""
    rng = np.random.default_rng(2)
    N, T, D = 1683, 32, 32
    ptr = (np.arange(2049) * 32).astype(np.uint64)
    ids = rng.integers(1, N, size=2048 * 32).astype(np.uint64)
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N)
    m = _build(pkg, "lstm", N, T, D, 4, "bpr", "adagrad", threads=256)
    l1 = m.fit(data)
    l2 = m.fit(data)
    assert np.isfinite(l1) and l2 < l1
    assert np.all(np.isfinite(m.get_parameter("item_embeddings")))


def test_shard_argument_checks(pkg):
    h = pkg.ewma.Hyperparameters(10, 4)
    with pytest.raises(pkg.SbrError):
        h.virtual_shards(3)
    with pytest.raises(pkg.SbrError):
        h.shard(2, 2)
    m = pkg.ewma.Hyperparameters(10, 4).embedding_dim(32).shard(0, 2).build()
    data = pkg.CompressedInteractions.from_csr([0, 4], [1, 2, 3, 4], None, num_items=10)
    with pytest.raises(pkg.SbrError):  # peers not attached yet
        m.fit(data)
    assert len(m.ipc_export()) == pkg.lib().sbr_model_ipc_handle_size()
