"""CPU-side tests of the product library: it loads, exports every symbol include/sbr_b200.h declares, its host
logic (CSR build, chunker) matches the oracle / the reference's golden vectors, and every compute entry point fails
LOUDLY without a GPU (there is no CPU fallback).  No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sbr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sbr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.lib()
    syms = header_symbols()
    assert len(syms) >= 45
    for s in syms:
        assert hasattr(lib, s), "libsbr_b200.so does not export %s" % s
    assert sorted(pkg.EXPORTS) == syms  # the Python mirror binds exactly the declared surface


def test_no_torch_in_abi_signatures():
    src = open(os.path.join(ROOT, "include", "sbr_b200.h")).read()
    assert "torch" not in src.lower() and "at::" not in src and "#include <cuda" not in src


def test_chunk_golden_vector_through_abi(pkg):
    """data.rs:630-662 through the product's chunker."""
    inter = pkg.Interactions.from_interactions([pkg.Interaction(0, i, i) for i in range(5)])
    assert inter.shape() == (1, 5)  # max + 1 (data.rs:202-203)
    c = inter.to_compressed()
    up, items, ts = c.arrays()
    got = [items[s:s + n].tolist() for s, n in c.user_chunks(0, 3)]
    assert got == [[0, 1], [2, 3, 4]]
    assert [ts[s:s + n].tolist() for s, n in c.user_chunks(0, 3)] == [[0, 1], [2, 3, 4]]
    with pytest.raises(pkg.SbrError):
        c.user_chunks(1, 3)  # get_user(user_id >= num_users) is None (data.rs:278-280)


def test_compress_matches_oracle_and_is_stable(pkg):
    rng = np.random.default_rng(0)
    for nu, ni, nnz in ((1, 1, 0), (5, 7, 1), (20, 20, 100), (300, 50, 5000)):
        users = rng.integers(0, nu, size=nnz).astype(np.uint64)
        items = rng.integers(0, ni, size=nnz).astype(np.uint64)
        ts = rng.integers(0, 10, size=nnz).astype(np.uint64)  # few distinct timestamps => many ties
        c = pkg.Interactions.from_arrays(users, items, ts, nu, ni).to_compressed()
        up, ii, tt = c.arrays()
        oup, oii, ott = O.compress(users, items, ts, nu)
        assert np.array_equal(up, oup) and np.array_equal(ii, oii) and np.array_equal(tt, ott)
        assert c.num_users() == nu and c.num_items() == ni and len(c) == nnz
        # round trip (data.rs:588-627): to_interactions() gives back the same multiset
        back = c.to_interactions()
        a = sorted(zip(users.tolist(), items.tolist(), ts.tolist()))
        b = sorted(zip(np.asarray(back._u).tolist(), np.asarray(back._i).tolist(), np.asarray(back._t).tolist()))
        assert a == b


def test_ml100k_through_product_csr(pkg, ml100k):
    c = pkg.Interactions.from_arrays(ml100k["raw_users"], ml100k["raw_items"], ml100k["raw_ts"]).to_compressed()
    assert c.shape() == (944, 1683)
    up, ii, tt = c.arrays()
    assert np.array_equal(up, ml100k["user_ptr"]) and np.array_equal(ii, ml100k["item_ids"])
    assert np.array_equal(tt, ml100k["timestamps"])
    # chunker agrees with the oracle on every user, at the three sequence lengths the configs use
    for T in (32, 128, 200):
        for u in range(0, 944, 37):
            assert c.user_chunks(u, T) == O.chunks(int(up[u + 1] - up[u]), T)


def test_invalid_arguments_are_status_codes(pkg):
    with pytest.raises(pkg.SbrError):  # user id >= num_users: Rust would panic on the index
        pkg.Interactions.from_arrays([5], [1], [0], num_users=3, num_items=4).to_compressed()
    with pytest.raises(pkg.SbrError):
        pkg.Interactions.from_arrays([1], [9], [0], num_users=3, num_items=4).to_compressed()
    with pytest.raises(pkg.SbrError):
        pkg.CompressedInteractions.from_csr([1, 2], [0, 0], None, num_items=3)  # ptr[0] != 0
    with pytest.raises(pkg.SbrError):
        pkg.CompressedInteractions.from_csr([0, 3, 2], [0, 0, 0], None, num_items=3)  # decreasing
    h = pkg.ewma.Hyperparameters(10, 5)
    with pytest.raises(pkg.SbrError):
        h.loss(7)
    with pytest.raises(pkg.SbrError):
        h._set("sbr_hyper_lstm_variant", 0)  # EWMA has no variant (ewma.rs:45-57)


def test_hyperparameter_builder_chains(pkg):
    """lstm.rs:54-138: every setter returns the builder."""
    h = (pkg.lstm.Hyperparameters(100, 32).learning_rate(0.16).l2_penalty(4e-4).embedding_dim(32).num_epochs(10)
         .loss(pkg.Loss.WARP).lstm_variant(pkg.LSTMVariant.Normal).num_threads(2)
         .parallelism(pkg.Parallelism.Synchronous).from_seed(bytes([42] * 16)).optimizer(pkg.Optimizer.Adagrad))
    assert isinstance(h, pkg.lstm.Hyperparameters)


def test_hyper_values_and_defaults(pkg):
    """Defaults of Hyperparameters::new (lstm.rs:56-71 / ewma.rs:61-76) read back through sbr_hyper_get_values."""
    v = pkg.lstm.Hyperparameters(100, 32).values()
    assert (v["model"], v["num_items"], v["max_sequence_length"], v["embedding_dim"]) == (0, 100, 32, 16)
    assert abs(v["learning_rate"] - 0.01) < 1e-9 and v["l2_penalty"] == 0.0
    assert (v["lstm_variant"], v["loss"], v["optimizer"], v["parallelism"]) == (
        pkg.LSTMVariant.Coupled, pkg.Loss.BPR, pkg.Optimizer.Adam, pkg.Parallelism.Synchronous)
    assert v["num_epochs"] == 10 and v["exact_arithmetic"] == 0
    h = pkg.ewma.Hyperparameters(7, 5).from_seed(bytes(range(16))).exact_arithmetic().num_threads(3)
    v = h.values()
    assert (v["model"], v["seed"], v["exact_arithmetic"], v["num_threads"]) == (1, bytes(range(16)), 1, 3)


@pytest.mark.parametrize("kind", ["lstm", "ewma"])
def test_hyperparameters_random(pkg, kind):
    """Hyperparameters::random (lstm.rs:141-172 / ewma.rs:139-170): every field inside the reference's ranges, the
    caller's rng advances, the same rng state gives the same draw, and over many draws every branch is taken."""
    H = pkg.lstm.Hyperparameters if kind == "lstm" else pkg.ewma.Hyperparameters
    st = (1, 2, 3, 4)
    seen = {"loss": set(), "optimizer": set(), "parallelism": set(), "lstm_variant": set(), "T": set(), "D": set(), "epochs": set()}
    first = None
    for i in range(200):
        h, st2 = H.random(1683, st)
        v = h.values()
        if first is None:
            first = v
            h_again, st_again = H.random(1683, st)
            va = h_again.values()
            assert st_again == st2 and {k: va[k] for k in va if k != "seed"} == {k: v[k] for k in v if k != "seed"}
        assert st2 != st
        st = st2
        assert v["num_items"] == 1683
        assert v["max_sequence_length"] in (16, 32, 64, 128) and v["embedding_dim"] in (16, 32, 64, 128)
        assert 1e-3 <= v["learning_rate"] < 10 ** 0.5 + 1e-6 and 1e-7 <= v["l2_penalty"] < 1e-3 + 1e-9
        assert v["loss"] in (pkg.Loss.BPR, pkg.Loss.Hinge) and v["optimizer"] in (0, 1) and v["parallelism"] in (0, 1)
        assert 1 <= v["num_threads"] <= (os.cpu_count() or 1) and v["num_epochs"] in (8, 16, 32, 64)
        for k, f in (("loss", "loss"), ("optimizer", "optimizer"), ("parallelism", "parallelism"), ("lstm_variant", "lstm_variant"),
                     ("T", "max_sequence_length"), ("D", "embedding_dim"), ("epochs", "num_epochs")):
            seen[k].add(v[f])
    assert seen["loss"] == {0, 1} and seen["optimizer"] == {0, 1} and seen["parallelism"] == {0, 1}
    assert seen["T"] == {16, 32, 64, 128} and seen["D"] == {16, 32, 64, 128} and seen["epochs"] == {8, 16, 32, 64}
    assert seen["lstm_variant"] == ({0, 1} if kind == "lstm" else {pkg.LSTMVariant.Coupled})
    with pytest.raises(pkg.SbrError):
        H.random(10, (0, 0, 0, 0))


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="this check is for GPU-less boxes")
def test_compute_fails_loudly_without_gpu(pkg):
    assert pkg.device_count() == 0
    with pytest.raises(pkg.SbrError) as e:
        pkg.lstm.Hyperparameters(10, 4).build()
    assert e.value.status == pkg.SBR_ERR_CUDA and "no CPU fallback" in str(e.value)
    c = pkg.Interactions.from_arrays([0, 0, 0], [1, 2, 3], [0, 1, 2]).to_compressed()
    with pytest.raises(pkg.SbrError) as e:
        c.upload()
    assert e.value.status == pkg.SBR_ERR_CUDA


def test_product_never_touches_the_oracle():
    """The product path must not route through oracle/: no source under the package or include/ mentions it."""
    for base in ("sbr-rs_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath:
                continue
            for f in files:
                if f.endswith((".cu", ".cuh", ".h", ".cc", ".py", "Makefile")):
                    txt = open(os.path.join(dirpath, f)).read()
                    code = re.sub(r"//.*|/\*.*?\*/|#.*", "", txt)  # comments may cite the oracle, code may not
                    assert "liboracle" not in code and "oracle_lib" not in code and "sbr_oracle.h" not in code, (dirpath, f)
                    assert not re.search(r"\bsbo_\w+\s*\(", code), (dirpath, f)
    # and the shared object has no undefined sbo_* symbols
    import subprocess
    out = subprocess.run(["nm", "-D", os.path.join(ROOT, "sbr-rs_b200", "libsbr_b200.so")], capture_output=True, text=True)
    assert "sbo_" not in out.stdout


@pytest.mark.parametrize("T,min_len,max_len", [(3, 0, 11), (7, 0, 40), (32, 1, 100), (32, 32, 32), (200, 0, 450)])
def test_host_schedule_equals_oracle(pkg, oracle, T, min_len, max_len):
    """The schedule fit() builds on the host -- chunks of every user with the first chunk the short one
    (data.rs:406-432), the len > 2 filter (sequence_model.rs:81), the master-rng Fisher-Yates shuffle (:84) -- is
    bit-identical to the oracle's restatement (sbo_subsequences + sbo_shuffle_u32), rng state afterwards included.
    (The library draws the swap partners 16 iterations ahead and chunks with one division per user.)"""
    import ctypes as C
    rng = np.random.default_rng(T * 1000 + max_len)
    U = 3000
    lens = rng.integers(min_len, max_len + 1, size=U)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ids = rng.integers(1, 50, size=int(ptr[-1])).astype(np.uint64)
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=50)
    seed = bytes(range(7, 23))
    r = oracle.make_rng(seed)
    state0 = (r.x, r.y, r.z, r.w)
    starts, slens, order, state1 = data.host_schedule(T, state0)
    ost, oln = oracle.subsequences(ptr, T)
    assert np.array_equal(starts, ost) and np.array_equal(slens, oln)
    oorder = np.arange(len(ost), dtype=np.uint32)
    oracle.lib().sbo_shuffle_u32(C.byref(r), oorder.ctypes.data_as(oracle.u32p), len(oorder))
    assert np.array_equal(order, oorder)
    assert state1 == (r.x, r.y, r.z, r.w)
    assert int(slens.min()) > 2 and int(slens.max()) <= T


@pytest.mark.parametrize("nsub,partitions,threads", [(1, 1, 4), (2, 2, 4), (1000, 7, 4), (140001, 128, 2), (300000, 0, 4), (600000, 4096, 8)])
def test_master_schedule_with_jump_ahead_equals_oracle(pkg, oracle, nsub, partitions, threads):
    """What fit() does with the master rng -- Fisher-Yates over the sub-sequence indices (sequence_model.rs:84), then one
    [u8; 16] seed per partition (:97) -- with the xorshift128 stream produced by several host threads (jump-ahead through
    the GF(2) transition matrix, partner list built in stream order with gen_range's redraws) is bit-identical to the
    oracle's plain loops and to the library's own one-thread path: order, partition keys, rng state afterwards."""
    import ctypes as C
    r = oracle.make_rng(bytes(range(3, 19)))
    state0 = (r.x, r.y, r.z, r.w)
    order, keys, state1 = pkg.host_master_schedule(state0, nsub, partitions, threads)
    order1, keys1, state11 = pkg.host_master_schedule(state0, nsub, partitions, 1)
    assert np.array_equal(order, order1) and np.array_equal(keys, keys1) and state1 == state11
    oorder = np.arange(nsub, dtype=np.uint32)
    L = oracle.lib()
    L.sbo_shuffle_u32(C.byref(r), oorder.ctypes.data_as(oracle.u32p), nsub)
    assert np.array_equal(order, oorder)
    okeys = []
    for _ in range(partitions):
        seed = bytes(L.sbo_rng_next_u32(C.byref(r)) & 0xFF for _ in range(16))
        okeys.append(int.from_bytes(seed[:8], "little"))
    assert np.array_equal(keys, np.array(okeys, dtype=np.uint64))
    assert state1 == (r.x, r.y, r.z, r.w)
    assert np.array_equal(np.sort(order), np.arange(nsub, dtype=np.uint32))


def test_threaded_subsequence_count_of_fit_equals_the_chunker(pkg, oracle):
    """fit() counts the sub-sequences with host threads (>= 2^18 users) and leaves the chunks to the device; sbr_host_schedule
    cross-checks that count against its own one-thread chunker, which is pinned on the oracle here at a size that takes the
    threaded path."""
    rng = np.random.default_rng(5)
    U, T = 300_000, 32
    lens = rng.integers(0, 70, size=U)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ids = rng.integers(1, 50, size=int(ptr[-1])).astype(np.uint64)
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=50)
    starts, slens, order, _ = data.host_schedule(T, (1, 2, 3, 4))
    ost, oln = oracle.subsequences(ptr, T)
    assert np.array_equal(starts, ost) and np.array_equal(slens, oln)
    assert np.array_equal(np.sort(order), np.arange(len(ost), dtype=np.uint32))


def test_host_schedule_empty_is_no_interactions(pkg):
    ptr = np.array([0, 2, 3, 5], dtype=np.uint64)          # every user has <= 2 interactions: nothing survives the filter
    ids = np.array([1, 2, 3, 4, 5], dtype=np.uint64)
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=10)
    with pytest.raises(pkg.NoInteractions):                # FittingError::NoInteractions (sequence_model.rs:86-88)
        data.host_schedule(8, (1, 2, 3, 4))


def test_user_based_split_equals_oracle_and_reference_recipe(pkg, oracle, ml100k):
    """data.rs:69-88 in the library (own SipHash-2-4) against the oracle's restatement at the reference's own test
    recipe: seed [42;16], test fraction 0.2 (lstm.rs:428-430); identical mask, identical rng state afterwards; no user
    on both sides."""
    up = ml100k["user_ptr"].astype(np.int64)
    users = np.repeat(np.arange(944), np.diff(up)).astype(np.uint64)
    r = oracle.make_rng(bytes([42] * 16))
    mask, state = pkg.user_based_split(users, (r.x, r.y, r.z, r.w), 0.2)
    omask, r2 = oracle.user_based_split(users, bytes([42] * 16), 0.2)
    assert np.array_equal(mask, omask) and state == (r2.x, r2.y, r2.z, r2.w)
    train_users, test_users = set(users[mask].tolist()), set(users[~mask].tolist())
    assert not (train_users & test_users) and 0.12 < len(test_users) / 943 < 0.28


def test_train_test_split_is_the_reference_shuffle(pkg, oracle):
    """data.rs:54-64: Fisher-Yates shuffle with the caller's rng, the first (fraction * len) as usize are the test set."""
    import ctypes as C
    n = 1000
    r = oracle.make_rng(bytes(range(16)))
    train, test, state = pkg.train_test_split(n, (r.x, r.y, r.z, r.w), 0.25)
    operm = np.arange(n, dtype=np.uint32)
    oracle.lib().sbo_shuffle_u32(C.byref(r), operm.ctypes.data_as(oracle.u32p), n)
    assert len(test) == 250 and len(train) == 750
    assert np.array_equal(np.concatenate([test, train]).astype(np.uint32), operm) and state == (r.x, r.y, r.z, r.w)
    assert sorted(np.concatenate([train, test]).tolist()) == list(range(n))
