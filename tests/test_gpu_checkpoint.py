"""Checkpoint / warm restart: the reference's model is Serialize/Deserialize as a whole -- parameters,
optimizer state inside HogwildParameter, the hyper-parameter rng (lstm.rs:204-210,386-389) -- and a second fit() call
continues from the current parameters and optimizer state.  save_state()/load_state() of the Python mirror write the
canonical host-order blobs behind sbr_model_get/set_parameter, the rng state and the update counter; a model restored
from the file continues bit-for-bit like the one that was saved (num_threads = 1: deterministic schedule)."""
import numpy as np
import pytest

from helpers import random_csr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,opt", [("ewma", "adagrad"), ("lstm", "adagrad"), ("lstm", "adam")])
def test_save_load_continues_bit_for_bit(pkg, tmp_path, kind, opt):
    rng = np.random.default_rng(11)
    N, T, D = 300, 12, 32
    ptr, ids = random_csr(rng, 60, N, 3, 30)

    def build():
        H = pkg.lstm.Hyperparameters if kind == "lstm" else pkg.ewma.Hyperparameters
        h = (H(N, T).embedding_dim(D).learning_rate(0.05).l2_penalty(1e-4).loss(pkg.Loss.WARP)
             .optimizer(pkg.Optimizer.Adam if opt == "adam" else pkg.Optimizer.Adagrad).num_epochs(1).num_threads(1)
             .from_seed(bytes(range(5, 21))))
        return (h.lstm_variant(pkg.LSTMVariant.Coupled) if kind == "lstm" else h).build()

    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N)
    a = build()
    a.fit(data)
    path = str(tmp_path / "model.npz")
    a.save_state(path)
    b = build()
    b.load_state(path)
    sa, sb = a.state_dict(), b.state_dict()
    assert set(sa) == set(sb) and any(k.endswith(".s1") for k in sa) and (opt != "adam" or any(k.endswith(".s2") for k in sa))
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    la, lb = a.fit(data), b.fit(data)           # warm restart: both continue from the same state
    assert la == lb
    for k, v in a.state_dict().items():
        assert np.array_equal(v, b.state_dict()[k]), k
    assert a.num_updates == b.num_updates and a.num_updates > 0
