"""Multi-process worker for the row-sharded model (one process per GPU, launched by torchrun or mp.spawn).

Checks, on every rank:  the table initialises identically to an unsharded model, remote rows gather bit-exactly through
the CUDA-IPC peer mappings, set/get_parameter round-trip across shards, concurrent fit() on disjoint users trains ONE
shared model (all ranks read back the same parameters, loss falls), and nothing goes non-finite.
Usage: torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tests/dist_worker.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402


def exchange_handles(model, world):
    """all-gather of the per-rank IPC handle blobs, rank order (host plumbing: any transport works)"""
    blobs = [None] * world
    dist.all_gather_object(blobs, model.ipc_export())
    return blobs


def split_users(ptr, ids, rank, world):
    """rank r trains the users u with u % world == r (disjoint, covers everything)"""
    lens = np.diff(ptr.astype(np.int64))
    mine = np.arange(len(lens)) % world == rank
    new_ptr = np.concatenate([[0], np.cumsum(lens[mine])]).astype(np.uint64)
    pieces = [ids[int(ptr[u]):int(ptr[u + 1])] for u in np.nonzero(mine)[0]]
    return new_ptr, (np.concatenate(pieces) if pieces else np.zeros(0, dtype=np.uint64))


def multi_rank_round_equals_oracle(pkg, rank, world, seed):
    """One synchronous round over ALL ranks' partitions against the oracle, element-wise: every partition of every rank trains
    one sequence over its own items; the round's result is: every visited row = the oracle's entries (gradients taken at the
    initial parameters) applied un-merged in the reference order -- rank-major, then partition, then the oracle's order inside
    a sequence -- whoever owns the row and on whichever GPU the sequence was computed; alpha = ONE Adagrad step on the
    gradient summed over the world * P sequences."""
    import ctypes as C
    import oracle_lib as O
    from test_gpu_lstm_tc import _adagrad
    N, T, D, P, lr, l2 = 200_003, 8, 32, 8, 0.05, 1e-3
    ptr = (np.arange(P + 1) * T).astype(np.uint64)
    all_ids = [(1000 + 5000 * r + np.arange(P * T)).astype(np.uint64) for r in range(world)]
    h = (pkg.ewma.Hyperparameters(N, T).embedding_dim(D).learning_rate(lr).l2_penalty(l2).loss(pkg.Loss.BPR)
         .optimizer(pkg.Optimizer.Adagrad).parallelism(pkg.Parallelism.Synchronous).num_epochs(1).num_threads(P).from_seed(seed))
    gm = h.shard(rank, world).build()
    gm.ipc_attach(exchange_handles(gm, world))
    r0 = np.random.default_rng(11)
    E0 = (r0.standard_normal((N, D)) * 0.3).astype(np.float32)
    b0 = (r0.standard_normal(N) * 0.3).astype(np.float32)
    A0 = (r0.standard_normal(D) * 0.5).astype(np.float32)
    dist.barrier()
    gm.set_parameter("item_embeddings", E0.ravel()); gm.set_parameter("item_biases", b0)   # every rank writes its own rows
    gm.set_parameter("alpha", A0)
    for n_ in ("item_embeddings", "item_biases", "alpha"):
        gm.set_parameter(n_ + ".s1", np.ones(len(gm.get_parameter(n_)), dtype=np.float32))
    dist.barrier()
    om = O.OracleModel("ewma", N, T, embedding_dim=D, learning_rate=lr, l2_penalty=l2, loss="bpr", optimizer="adagrad",
                       parallelism="synchronous", num_threads=P, num_epochs=1, seed=seed)
    om.param("item_embeddings")[:] = E0.ravel(); om.param("item_biases")[:] = b0; om.param("alpha")[:] = A0
    for n_ in ("item_embeddings", "item_biases", "alpha"):
        om.param(n_ + ".s1")[:] = 1.0
    L = O.lib()
    rng = O.Rng(*gm.rng_state)                       # the same master rng on every rank: same shuffle, same partition keys
    order = np.arange(P, dtype=np.uint32)
    L.sbo_shuffle_u32(C.byref(rng), order.ctypes.data_as(O.u32p), P)
    keys = []
    for _ in range(P):
        sd = bytes(L.sbo_rng_next_u32(C.byref(rng)) & 0xFF for _ in range(16))
        keys.append(int.from_bytes(sd[:8], "little"))
    E, SE, b, Sb = E0.copy(), np.ones_like(E0), b0.copy(), np.ones_like(b0)
    dsum = np.zeros(D, dtype=np.float32)
    touched = {}
    for r in range(world):                           # rank-major = the engine's application order (part_base = rank * P)
        for p in range(P):
            sq = int(order[p])
            seq = all_ids[r][sq * T:(sq + 1) * T]
            _, negs, dg = om.step(seq, key=keys[p], step=0, apply=False)
            dsum += dg
            rows, grads, brows, bgrads = om.last_sparse_grads()
            for row in set(rows.tolist()) | set(brows.tolist()):
                touched.setdefault(row, set()).add((r, p))
            for row, gr in zip(rows.tolist(), grads):
                E[row], SE[row] = _adagrad(E[row], SE[row], gr, lr, l2)
            for row, gr in zip(brows.tolist(), bgrads.tolist()):
                wv, gv = _adagrad(b[row:row + 1], Sb[row:row + 1], gr, lr, l2)
                b[row], Sb[row] = wv[0], gv[0]
    A1, _ = _adagrad(A0, np.ones_like(A0), dsum, lr, l2)
    data = pkg.CompressedInteractions.from_csr(ptr, all_ids[rank], None, num_items=N)
    dist.barrier()
    gm.fit(data)
    dist.barrier()
    # (partition p draws the same negatives on every rank -- same key, same step -- so a negative's row is visited once per
    # rank: the expected values above took those entries in the engine's order, rank-major; nothing needs to be left out)
    clean = np.array(sorted(touched), dtype=np.int64)
    assert sum(1 for v in touched.values() if len(v) > 1) >= P
    gE = gm.get_parameter("item_embeddings").reshape(N, D)
    gb = gm.get_parameter("item_biases")
    gS = gm.get_parameter("item_embeddings.s1").reshape(N, D)
    assert np.abs(E[clean] - E0[clean]).mean() > 1e-4
    assert np.abs(gE[clean] - E[clean]).max() <= 2e-4, np.abs(gE[clean] - E[clean]).max()
    assert np.abs(gb[clean] - b[clean]).max() <= 2e-4
    assert np.abs(gS[clean] - SE[clean]).max() <= 2e-4
    assert np.abs(gm.get_parameter("alpha") - A1).max() <= 2e-4, np.abs(gm.get_parameter("alpha") - A1).max()
    untouched = np.setdiff1d(np.arange(N), np.array(sorted(touched), dtype=np.int64))
    assert np.array_equal(gE[untouched], E0[untouched])
    assert gm.last_fit_stats()["partitions"] == P
    if rank == 0:
        print("dist_worker sync oracle-equality OK world=%d rows=%d kernel=%s" % (world, len(clean), gm.last_fit_stats()["kernel"]))
    dist.barrier()
    del gm


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    pkg = g.load_package()
    on_gpu = pkg.device_count() > 0
    dist.init_process_group("gloo")
    rng = np.random.default_rng(0)
    N, T, D, U = 1001, 16, 32, 512
    lens = rng.integers(3, 40, size=U)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ids = rng.integers(1, N, size=int(ptr[-1])).astype(np.uint64)
    my_ptr, my_ids = split_users(ptr, ids, rank, world)
    # host logic that needs no GPU: the split is a partition of the users
    counts = [None] * world
    dist.all_gather_object(counts, (len(my_ptr) - 1, int(my_ptr[-1])))
    assert sum(c[0] for c in counts) == U and sum(c[1] for c in counts) == int(ptr[-1])
    # bench.py's replicated multi-GPU mode: after a step every rank holds start + the SUM of all ranks' deltas
    import bench

    class FakeModel:
        def __init__(self):
            self.p = {"a": np.arange(6, dtype=np.float32), "b": np.ones(3, dtype=np.float32)}
        def get_parameter(self, n):
            return self.p[n].copy()
        def set_parameter(self, n, v):
            self.p[n] = np.asarray(v, dtype=np.float32).copy()
    if not on_gpu:
        fm = FakeModel()
        rs = bench.ReplicaSync(fm, ["a", "b"], torch, dist)
        fm.p["a"] = fm.p["a"] + np.float32(rank + 1)          # this rank's local training moved the parameters
        fm.p["b"] = fm.p["b"] * np.float32(2 + rank)
        rs()
        tot = sum(r + 1 for r in range(world))
        assert np.allclose(fm.p["a"], np.arange(6) + tot) and np.allclose(fm.p["b"], 1 + sum(1 + r for r in range(world)))
        assert rs.bytes_per_sync == 36
        blobs = [None] * world
        dist.all_gather_object(blobs, bytes([rank]) * 192)
        assert [b[0] for b in blobs] == list(range(world))
        dist.barrier()
        if rank == 0:
            print("dist_worker host-only OK world=%d" % world)
        return

    torch.cuda.set_device(local)
    pkg.set_device(local)
    seed = bytes(range(16))
    for kind in ("ewma", "lstm"):
        H = pkg.lstm.Hyperparameters if kind == "lstm" else pkg.ewma.Hyperparameters
        def hyper():
            h = H(N, T).embedding_dim(D).learning_rate(0.05).l2_penalty(1e-4).loss(pkg.Loss.BPR) \
                .optimizer(pkg.Optimizer.Adagrad).parallelism(pkg.Parallelism.Asynchronous).num_epochs(1).num_threads(8) \
                .from_seed(seed)
            return h.lstm_variant(pkg.LSTMVariant.Normal) if kind == "lstm" else h
        ref = hyper().build()                                   # unsharded twin, local
        model = hyper().shard(rank, world).build()
        model.ipc_attach(exchange_handles(model, world))
        dist.barrier()
        names = ["item_embeddings", "item_biases"] + (["lstm_weights", "lstm_biases"] if kind == "lstm" else ["alpha"])
        for n in names:
            assert np.array_equal(ref.get_parameter(n), model.get_parameter(n)), (kind, n)
        probe = rng.integers(0, N, size=777).astype(np.uint64)
        assert np.array_equal(ref.gather_rows(probe).view(np.uint32), model.gather_rows(probe).view(np.uint32))
        e = np.random.default_rng(7).standard_normal(N * D).astype(np.float32) * 0.1
        dist.barrier()
        model.set_parameter("item_embeddings", e)               # every rank writes its own rows
        dist.barrier()
        assert np.array_equal(model.get_parameter("item_embeddings"), e)
        data = pkg.CompressedInteractions.from_csr(my_ptr, my_ids, None, num_items=N)
        losses = []
        for _ in range(4):
            dist.barrier()
            losses.append(model.fit(data) / 8)                  # all ranks train the one shared model concurrently
        dist.barrier()
        after = model.get_parameter("item_embeddings")
        assert np.all(np.isfinite(after)) and np.abs(after - e).max() > 1e-3
        mine = torch.tensor(after[:4096].astype(np.float64))
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        for t in gathered:                                      # every rank sees the same table
            assert torch.equal(t, gathered[0])
        assert losses[-1] < losses[0], losses
        mrr = pkg.mrr_score(model, data)
        assert 0.0 < mrr <= 1.0
        if rank == 0:
            print("dist_worker %s OK world=%d losses=%s" % (kind, world, [round(x, 4) for x in losses]))
        dist.barrier()
        del model, ref
    # ---- Parallelism::Synchronous across GPUs: NCCL all-to-all exchange of ids / rows / gradient rows (sync_engine.cu) ----
    uid = [pkg.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    pkg.dist_init(rank, world, uid[0])
    # ---- replicas: sbr_model_replica_sync = start + the SUM of all ranks' deltas, parameters and optimizer state, on device ----
    rm = (pkg.lstm.Hyperparameters(N, T).embedding_dim(D).optimizer(pkg.Optimizer.Adagrad).parallelism(pkg.Parallelism.Asynchronous)
          .num_threads(8).from_seed(seed).build())
    assert rm.replica_sync() == 0                                  # first call: records the common starting point
    start = {n: rm.get_parameter(n) for n in ("item_embeddings", "item_biases", "lstm_weights", "lstm_biases", "item_embeddings.s1")}
    for n, v in start.items():                                      # this rank's "training": a rank-dependent change of every blob
        rm.set_parameter(n, v + np.float32(0.25 * (rank + 1)) * np.sign(v + 1e-3).astype(np.float32))
    nbytes = rm.replica_sync()
    assert nbytes == 4 * (N * (4 + 2 * D) + 3 * (2 * D * 4 * D + 4 * D)), nbytes
    tot = np.float32(sum(0.25 * (r + 1) for r in range(world)))
    for n, v in start.items():
        want = v + tot * np.sign(v + 1e-3).astype(np.float32)
        assert np.allclose(rm.get_parameter(n), want, atol=1e-5), n
    assert rm.replica_sync() == nbytes                              # nothing changed since: the replicas stay put
    for n, v in start.items():
        assert np.allclose(rm.get_parameter(n), v + tot * np.sign(v + 1e-3).astype(np.float32), atol=1e-5), n
    if rank == 0:
        print("dist_worker replica sync OK world=%d bytes=%d" % (world, nbytes))
    del rm
    Nbig = 300_001
    ids_big = np.random.default_rng(3).integers(1, Nbig, size=int(ptr[-1])).astype(np.uint64)
    my_ptr2, my_ids2 = split_users(ptr, ids_big, rank, world)
    hs = (pkg.ewma.Hyperparameters(Nbig, T).embedding_dim(D).learning_rate(0.05).l2_penalty(1e-4).loss(pkg.Loss.BPR)
          .optimizer(pkg.Optimizer.Adagrad).parallelism(pkg.Parallelism.Synchronous).num_epochs(1).num_threads(8)
          .from_seed(seed).shard(rank, world))
    sm = hs.build()
    sm.ipc_attach(exchange_handles(sm, world))   # only so that get_parameter can read every shard; fit() does not use it
    e0 = sm.get_parameter("item_embeddings").copy()
    data2 = pkg.CompressedInteractions.from_csr(my_ptr2, my_ids2, None, num_items=Nbig)
    sl = []
    for _ in range(4):
        dist.barrier()
        sl.append(sm.fit(data2) / 8)
    dist.barrier()
    e1 = sm.get_parameter("item_embeddings")
    al = torch.tensor(sm.get_parameter("alpha").astype(np.float64))
    als = [torch.zeros_like(al) for _ in range(world)]
    dist.all_gather(als, al)
    for t in als:
        assert torch.equal(t, als[0])                            # dense replicas stay bit-identical
    assert np.all(np.isfinite(e1)) and np.abs(e1 - e0).max() > 1e-3 and sl[-1] < sl[0], sl
    chk = torch.tensor([float(np.abs(e1).sum())], dtype=torch.float64)
    chks = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(chks, chk)
    assert all(torch.equal(c, chks[0]) for c in chks)           # every rank reads the same table
    st = sm.last_fit_stats()
    assert st["kernel_launches"] > 10
    multi_rank_round_equals_oracle(pkg, rank, world, seed)
    if rank == 0:
        print("dist_worker sync OK world=%d losses=%s launches=%d" % (world, [round(x, 4) for x in sl], st["kernel_launches"]))
    dist.barrier()
    del sm
    pkg.dist_finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
