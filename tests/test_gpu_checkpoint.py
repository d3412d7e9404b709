"""Checkpoint / warm restart: the reference's model is Serialize/Deserialize as a whole -- hyperparameters,
parameters, optimizer state inside HogwildParameter, the hyper-parameter rng (lstm.rs:38-52,204-210,386-389) -- and a
second fit() call continues from the current parameters and optimizer state.  sbr_model_save / sbr_model_load /
sbr_model_restore of the C library write and read ONE flat file (layout documented in include/sbr_b200.h); a model
loaded from the file continues bit-for-bit like the one that was saved (num_threads = 1: deterministic schedule).
Everything here goes through the C ABI (ctypes); the file header is additionally parsed by hand against the documented
layout."""
import struct

import numpy as np
import pytest

from helpers import random_csr

pytestmark = pytest.mark.gpu


def _names(kind, opt):
    base = ["item_embeddings", "item_biases"] + (["lstm_weights", "lstm_biases"] if kind == "lstm" else ["alpha"])
    out = []
    for b in base:
        out += [b, b + ".s1"] + ([b + ".s2"] if opt == "adam" else [])
    return out


@pytest.mark.parametrize("kind,opt", [("ewma", "adagrad"), ("lstm", "adagrad"), ("lstm", "adam")])
def test_save_load_continues_bit_for_bit(pkg, tmp_path, kind, opt):
    rng = np.random.default_rng(11)
    N, T, D = 300, 12, 32
    ptr, ids = random_csr(rng, 60, N, 3, 30)
    seed = bytes(range(5, 21))
    H = pkg.lstm.Hyperparameters if kind == "lstm" else pkg.ewma.Hyperparameters
    h = (H(N, T).embedding_dim(D).learning_rate(0.05).l2_penalty(1e-4).loss(pkg.Loss.WARP)
         .optimizer(pkg.Optimizer.Adam if opt == "adam" else pkg.Optimizer.Adagrad).num_epochs(1).num_threads(1)
         .from_seed(seed))
    a = (h.lstm_variant(pkg.LSTMVariant.Coupled) if kind == "lstm" else h).build()
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N)
    a.fit(data)
    path = str(tmp_path / "model.sbr")
    a.save(path)

    # the documented layout, parsed by hand
    raw = open(path, "rb").read()
    assert raw[:8] == b"SBRB200\0"
    version, header_bytes = struct.unpack_from("<II", raw, 8)
    assert version == 1 and header_bytes % 64 == 0
    model, variant, loss, optimizer, par, exact = struct.unpack_from("<6i", raw, 16)
    num_items, max_len, dim, threads, epochs = struct.unpack_from("<5Q", raw, 40)
    lr, l2 = struct.unpack_from("<2f", raw, 80)
    assert (model, loss, optimizer) == (0 if kind == "lstm" else 1, 2, 1 if opt == "adam" else 0)
    assert (num_items, max_len, dim, threads, epochs) == (N, T, D, 1, 1)
    assert abs(lr - 0.05) < 1e-9 and abs(l2 - 1e-4) < 1e-9 and raw[88:104] == seed
    assert struct.unpack_from("<4I", raw, 104) == tuple(a.rng_state)
    assert struct.unpack_from("<Q", raw, 120)[0] == a.num_updates > 0
    n_blobs = struct.unpack_from("<I", raw, 128)[0]
    names = _names(kind, opt)
    assert n_blobs == len(names)
    off_expect = header_bytes
    for i, n in enumerate(names):
        nm, ln, off = struct.unpack_from("<32sQQ", raw, 136 + 48 * i)
        assert nm.rstrip(b"\0").decode() == n and off == off_expect
        blob = np.frombuffer(raw, dtype="<f4", count=ln, offset=off)
        assert np.array_equal(blob, a.get_parameter(n)), n
        off_expect += 4 * ln
    assert off_expect == len(raw)

    b = pkg.load_model(path)                     # hyperparameters come from the file
    hv = b.hyper_values()
    assert hv == a.hyper_values() and hv["max_sequence_length"] == T and hv["seed"] == seed
    for n in names:
        assert np.array_equal(a.get_parameter(n), b.get_parameter(n)), n
    assert tuple(a.rng_state) == tuple(b.rng_state) and a.num_updates == b.num_updates
    la, lb = a.fit(data), b.fit(data)            # warm restart: both continue from the same state
    assert la == lb
    for n in names:
        assert np.array_equal(a.get_parameter(n), b.get_parameter(n)), n
    assert a.num_updates == b.num_updates

    c = (H(N, T).embedding_dim(D).optimizer(pkg.Optimizer.Adam if opt == "adam" else pkg.Optimizer.Adagrad)).build()
    c.restore(path)                              # restore into an existing model of the same shape
    for n in names:
        assert np.array_equal(np.frombuffer(raw, dtype="<f4", count=len(c.get_parameter(n)),
                                            offset=struct.unpack_from("<32sQQ", raw, 136 + 48 * names.index(n))[2]),
                              c.get_parameter(n)), n
    wrong = (H(N + 1, T).embedding_dim(D).optimizer(pkg.Optimizer.Adam if opt == "adam" else pkg.Optimizer.Adagrad)).build()
    with pytest.raises(pkg.SbrError):
        wrong.restore(path)


def test_load_rejects_garbage(pkg, tmp_path):
    p = tmp_path / "junk.sbr"
    p.write_bytes(b"not a checkpoint" * 20)
    with pytest.raises(pkg.SbrError):
        pkg.load_model(str(p))
    with pytest.raises(pkg.SbrError):
        pkg.load_model(str(tmp_path / "missing.sbr"))
