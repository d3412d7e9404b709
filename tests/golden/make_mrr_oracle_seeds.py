#!/usr/bin/env python
"""Golden oracle arm of tests/test_gpu_mrr.py: test MRR of the CPU oracle at BASELINE configs C1 / C2 over model seeds.

ML-100K (tests/golden/ml100k_csr.npz), user_based_split(0.2) with XorShift [42;16] (lstm.rs:428-430), seq 32, dim 32,
WARP, Adagrad lr 0.16 l2 4e-4, LSTMVariant::Normal, 10 epochs, num_threads 1 (deterministic).  Seed s fixes the initial
parameters (numpy Generator(1000 + s): embeddings N(0, 1/D) as at lstm.rs:22-25, LSTM weights U(+-1/sqrt(D)), biases /
alpha zero) and the model rng (from_seed(bytes([s + 1] * 16))), exactly as the GPU test sets them, so every arm of the
GPU test starts from the parameters the oracle started from.  Run here (CPU container), output committed:

    python tests/golden/make_mrr_oracle_seeds.py --seeds 64
"""
import argparse
import json
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
N, T, D, LR, L2, EPOCHS = 1683, 32, 32, 0.16, 4e-4, 10


def initial_parameters(kind, s):
    """shared with tests/test_gpu_mrr.py"""
    r = np.random.default_rng(1000 + s)
    p = {"item_embeddings": (r.standard_normal(N * D) / D).astype(np.float32), "item_biases": np.zeros(N, np.float32)}
    if kind == "lstm":
        a = 1.0 / np.sqrt(D)
        p["lstm_weights"] = r.uniform(-a, a, 2 * D * 4 * D).astype(np.float32)
        p["lstm_biases"] = r.uniform(-a, a, 4 * D).astype(np.float32)
    else:
        p["alpha"] = np.zeros(D, np.float32)
    return p


def split(O):
    z = np.load(os.path.join(HERE, "ml100k_csr.npz"))
    up = z["user_ptr"].astype(np.int64)
    users = np.repeat(np.arange(944), np.diff(up)).astype(np.uint64)
    items, ts = z["item_ids"].astype(np.uint64), z["timestamps"].astype(np.uint64)
    is_train, _ = O.user_based_split(users, bytes([42] * 16), 0.2)
    return (O.compress(users[is_train], items[is_train], ts[is_train], 944),
            O.compress(users[~is_train], items[~is_train], ts[~is_train], 944))


def one(job):
    import oracle_lib as O
    kind, s, threads, par = job
    tr, te = split(O)
    m = O.OracleModel(kind, N, T, embedding_dim=D, learning_rate=LR, l2_penalty=L2, lstm_variant="normal", loss="warp",
                      optimizer="adagrad", parallelism=par, num_threads=threads, num_epochs=EPOCHS,
                      seed=bytes([s + 1] * 16))
    for k, v in initial_parameters(kind, s).items():
        m.param(k)[:] = v
    assert m.fit(tr[0], tr[1])[0] == 0
    rc, mrr = m.mrr_score(te[0], te[1])
    assert rc == 0
    return kind, s, float(mrr)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=64)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--threads", type=int, default=1, help="> 1: the oracle's Hogwild mode (lock-free pthreads; NOT deterministic); "
                    "results are added to the JSON under <kind>_hogwild<threads>")
    ap.add_argument("--synchronous", action="store_true", help="with --threads > 1: Parallelism::Synchronous (the oracle's barrier mode, "
                    "deterministic; the reference default, lstm.rs:66); results under <kind>_sync<threads>")
    a = ap.parse_args()
    import oracle_lib as O
    O.lib()
    par = "synchronous" if a.synchronous else "asynchronous"
    with mp.Pool(a.procs) as pool:
        res = pool.map(one, [(k, s, a.threads, par) for k in ("lstm", "ewma") for s in range(a.seeds)], chunksize=1)
    path = os.path.join(HERE, "mrr_oracle_seeds.json")
    out = json.load(open(path)) if (a.threads > 1 and os.path.exists(path)) else {}
    out["recipe"] = "ML-100K split [42;16] 0.2, seq %d dim %d WARP Adagrad lr %g l2 %g Normal, %d epochs, 1 thread" % (T, D, LR, L2, EPOCHS)
    sfx = "" if a.threads == 1 else ("_sync%d" if a.synchronous else "_hogwild%d") % a.threads
    for k in ("lstm", "ewma"):
        out[k + sfx] = [m for kk, s, m in sorted(res, key=lambda r: (r[0], r[1])) if kk == k]
        print(k + sfx, "mean %.4f sd %.4f n %d" % (np.mean(out[k + sfx]), np.std(out[k + sfx], ddof=1), len(out[k + sfx])))
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
