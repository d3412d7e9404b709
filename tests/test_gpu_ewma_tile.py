"""EWMA tile kernel (kernels_ewma_tile.cu; num_threads a multiple of 128, D = 32) against the CPU oracle.

Same construction as tests/test_gpu_lstm_tc.py: every partition trains exactly one sequence over its own items, so
the Hogwild races between partitions cannot matter and the oracle's per-sequence gradients (taken at the initial
parameters) can be replayed element-wise.  The kernel's forward arithmetic is fp32 (packed f32x2 FMAs, MUFU
ex2/rcp/rsqrt); the saved activations (s_t, x_t, g (q - p)) are bf16 copies, so gradients carry ~0.4 % relative error:
stated tolerance 4e-4 on parameters at lr 0.05 (updates are O(1e-2)).
Dense parameter: every sequence applies its own optimizer step on alpha (one step per sub-sequence, ewma.rs:302 /
sequence_model.rs:163-169) as reduce-adds of the deltas; with every partition starting from the same alpha the result
is alpha0 + sum_p delta_p(alpha0).
"""
import ctypes as C

import numpy as np
import pytest

from helpers import make_pair, stream_csr
from test_gpu_lstm_tc import _adagrad, _adam

pytestmark = pytest.mark.gpu

CASES = [(128, "bpr", "adagrad"), (256, "bpr", "adagrad"), (256, "hinge", "adagrad"), (256, "warp", "adagrad"),
         (128, "warp", "adagrad"), (128, "bpr", "adam"), (128, "warp", "adam")]
DECISION_BAND = 2e-3


@pytest.mark.parametrize("P,loss,optimizer", CASES)
def test_one_round_matches_oracle_gradients(pkg, oracle, P, loss, optimizer):
    N, T, D, lr, l2 = 60000, 8, 32, 0.05, 1e-3
    adam = optimizer == "adam"
    ptr = (np.arange(P + 1) * T).astype(np.uint64)
    ids = (1000 + np.arange(P * T)).astype(np.uint64)
    gm, om = make_pair(pkg, oracle, "ewma", N, T, D, loss=loss, optimizer=optimizer, lr=lr, l2=l2, epochs=1, threads=P,
                       scale=0.3)
    rs = np.random.default_rng(9)
    if loss == "warp":
        gm.set_parameter("item_biases", rs.standard_normal(N).astype(np.float32))
    slot = ".s2" if adam else ".s1"
    for n in ("item_embeddings", "item_biases", "alpha"):
        gm.set_parameter(n + slot, np.ones(len(gm.get_parameter(n)), dtype=np.float32))
    for n in om.param_names():
        om.param(n)[:] = gm.get_parameter(n)
        om.param(n + slot)[:] = 1.0
    E0, b0 = om.param("item_embeddings").reshape(N, D).copy(), om.param("item_biases").copy()
    A0 = om.param("alpha").copy()

    L = oracle.lib()
    rng = oracle.Rng(*gm.rng_state)
    order = np.arange(P, dtype=np.uint32)
    L.sbo_shuffle_u32(C.byref(rng), order.ctypes.data_as(oracle.u32p), P)
    keys = []
    for _ in range(P):
        seed = bytes(L.sbo_rng_next_u32(C.byref(rng)) & 0xFF for _ in range(16))
        keys.append(int.from_bytes(seed[:8], "little"))

    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N)
    gm.fit(data)
    st = gm.last_fit_stats()
    assert st["partitions"] == P and st["kernel_launches"] == 1
    assert "ewma_tile" in st.get("kernel", "ewma_tile")

    E, SE1, SE2 = E0.copy(), (np.zeros_like(E0) if adam else np.ones_like(E0)), np.ones_like(E0)
    b, Sb1, Sb2 = b0.copy(), (np.zeros_like(b0) if adam else np.ones_like(b0)), np.ones_like(b0)
    A1 = A0.copy()
    touched, shaky = {}, set()
    tries_hist = np.zeros(6, dtype=np.int64)
    for p in range(P):
        sq = int(order[p])
        seq = ids[sq * T:(sq + 1) * T]
        _, negs, dg = om.step(seq, key=keys[p], step=0, apply=False)
        rows, grads, brows, bgrads = om.last_sparse_grads()
        for rrow in set(rows.tolist()) | set(brows.tolist()):
            touched.setdefault(rrow, set()).add(p)
        if loss != "bpr":
            for t in range(T - 1):
                h = om.user_representation(seq[:t + 1])[1]
                pos = float(h @ E0[int(seq[t + 1])] + b0[int(seq[t + 1])])
                chosen = None
                for j in range(5 if loss == "warp" else 1):
                    cand = int(L.sbo_draw_item(keys[p], 0, t, j, N))
                    margin = 1.0 - pos + float(h @ E0[cand] + b0[cand])
                    if abs(margin) < DECISION_BAND:
                        shaky.add(p)
                    chosen = cand
                    if margin > 0.0:
                        break
                assert chosen == int(negs[t])
                tries_hist[j + 1] += 1
        tstep = p + 1
        if adam:
            w1, _, _ = _adam(A0, np.zeros_like(A0), np.ones_like(A0), dg, lr, l2, tstep)
        else:
            w1, _ = _adagrad(A0, np.ones_like(A0), dg, lr, l2)
        if p not in shaky:
            A1 += w1 - A0
        nt_ = len(rows) // 3
        eorder = [3 * k for k in reversed(range(nt_))] + [3 * k + j for k in range(nt_) for j in (1, 2)]
        border = [2 * k for k in reversed(range(nt_))] + [2 * k + 1 for k in range(nt_)]
        rows, grads, brows, bgrads = rows[eorder], grads[eorder], brows[border], bgrads[border]
        for rrow, gr in zip(rows.tolist(), grads):
            if adam:
                E[rrow], SE1[rrow], SE2[rrow] = _adam(E[rrow], SE1[rrow], SE2[rrow], gr, lr, l2, tstep)
            else:
                E[rrow], SE1[rrow] = _adagrad(E[rrow], SE1[rrow], gr, lr, l2)
        for rrow, gr in zip(brows.tolist(), bgrads.tolist()):
            if adam:
                wv, mv, vv = _adam(b[rrow:rrow + 1], Sb1[rrow:rrow + 1], Sb2[rrow:rrow + 1], gr, lr, l2, tstep)
                b[rrow], Sb1[rrow], Sb2[rrow] = wv[0], mv[0], vv[0]
            else:
                wv, gv = _adagrad(b[rrow:rrow + 1], Sb1[rrow:rrow + 1], gr, lr, l2)
                b[rrow], Sb1[rrow] = wv[0], gv[0]
    if loss == "warp":
        assert tries_hist[2:].sum() > 0.03 * tries_hist.sum(), tries_hist
    assert len(shaky) <= 0.06 * P, (len(shaky), P)

    clean = np.array(sorted(k for k, v in touched.items() if len(v) == 1 and not (v & shaky)), dtype=np.int64)
    assert len(clean) > 0.85 * len(touched)
    gE = gm.get_parameter("item_embeddings").reshape(N, D)
    gb = gm.get_parameter("item_biases")
    gA = gm.get_parameter("alpha")
    assert np.all(np.isfinite(gE)) and np.all(np.isfinite(gb)) and np.all(np.isfinite(gA))
    # bf16 copies of s_t, x_t and g (q - p): every gradient term carries a relative rounding of 2^-9, an Adagrad step moves
    # an element by <= lr * |error of its gradient|.  BPR (g <= 0.25): 4e-4; hinge / WARP (g = 1, four times the gradients,
    # dx sums up to T - 1 rounded terms of size ~1): 1.5e-3 -- the same relative tolerance (~1 % of the largest updates)
    tol = 4e-4 if loss == "bpr" else 1.5e-3
    assert np.abs(E[clean] - E0[clean]).mean() > (2e-5 if adam else 3e-4)
    assert np.abs(gE[clean] - E[clean]).max() <= tol, np.abs(gE[clean] - E[clean]).max()
    assert np.abs(gb[clean] - b[clean]).max() <= tol
    g1 = gm.get_parameter("item_embeddings.s1").reshape(N, D)
    if adam:
        assert np.abs(g1[clean] - SE1[clean]).max() <= tol
        g2 = gm.get_parameter("item_embeddings.s2").reshape(N, D)
        assert np.abs(g2[clean] - SE2[clean]).max() <= tol
    else:
        # G = 1 + sum g^2 with every g off by <= delta (absolute: cancelling bf16 terms): |dG| <= 2 |g| delta + delta^2
        delta = 0.004 if loss == "bpr" else 0.016
        viol = np.abs(g1[clean] - SE1[clean]) - (3.0 * delta * np.sqrt(SE1[clean] - 1.0) + delta * delta + 1e-5)
        assert viol.max() <= 0, viol.max()
    dA = A1 - A0                      # every partition's step taken from (alpha0, state0), summed
    dG = gA - A0
    if not shaky and not adam:
        assert np.abs(dA).mean() > 1e-3, np.abs(dA).mean()
        if loss == "bpr":
            # small gradients (g'^2 << G = 1): the accumulator a step sees hardly depends on which partitions stepped before
            assert np.abs(dG - dA).max() <= 0.05 * np.abs(dA).max() + 2e-4, (np.abs(dG - dA).max(), np.abs(dA).max())
        else:
            # g = 1: partitions that reach their dense step later see the earlier ones' g'^2 in the accumulator (the
            # reduce-adds have landed) and take smaller steps -- anything between the sum from the common start and a
            # sequential pass is legitimate: same direction, comparable size
            big = np.abs(dA) > 0.25 * np.abs(dA).max()
            assert np.all(np.sign(dG[big]) == np.sign(dA[big]))
            assert 0.3 * np.abs(dA).mean() < np.abs(dG).mean() < 1.5 * np.abs(dA).mean()
    if adam:
        # Adam's dense step is a plain read-modify-write per sub-sequence (Hogwild, lost updates between concurrent partitions as
        # in the reference): alpha moved by at least one and at most all of the partitions' steps
        assert np.abs(dG).max() > 1e-5 and np.all(np.abs(dG) <= np.abs(dA) + lr * 1.01)
    untouched = np.setdiff1d(np.arange(N), np.array(sorted(touched), dtype=np.int64))
    assert np.array_equal(gE[untouched], E0[untouched])


@pytest.mark.parametrize("loss,optimizer,lr", [("bpr", "adagrad", 0.05), ("warp", "adagrad", 0.05), ("hinge", "adam", 0.002)])
def test_tile_kernel_learns_like_the_exact_path(pkg, oracle, loss, optimizer, lr):
    """Per-epoch losses of the tile kernel (256 partitions) track the exact warp-per-partition kernel run with the same
    partition count on an ML-100K-shaped stream."""
    rng = np.random.default_rng(5)
    N, T, D = 1683, 32, 32
    ptr, ids = stream_csr(rng, 16384, N, 32)
    losses = {}
    for kern in ("exact", "tile"):
        gm, _ = make_pair(pkg, oracle, "ewma", N, T, D, loss=loss, optimizer=optimizer, lr=lr, l2=1e-4, epochs=1,
                          threads=256, exact=(kern == "exact"))
        data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N)
        losses[kern] = [gm.fit(data) / 256 for _ in range(6)]
        assert ("ewma_tile" in gm.last_fit_stats().get("kernel", "")) == (kern == "tile")
        for n in ("item_embeddings", "item_biases", "alpha"):
            assert np.all(np.isfinite(gm.get_parameter(n))), (kern, n)
    a, b = np.array(losses["exact"]), np.array(losses["tile"])
    assert b[-1] < b[0]
    assert np.max(np.abs(a - b)) < (0.02 if optimizer == "adagrad" else 0.08), (a, b)
