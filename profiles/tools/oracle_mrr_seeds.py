#!/usr/bin/env python
"""Test MRR of the CPU oracle at the reference's own acceptance recipes, over model seeds.

The reference pins its LSTM / EWMA models with MRR floors on ML-100K (lstm.rs:450-520, ewma.rs:463-507): split
user_based_split(0.2) with XorShift [42;16] (lstm.rs:428-430), max_sequence_length 128, dim 32, lr 0.16, l2 4e-4,
LSTMVariant::Normal, Adagrad, 10 epochs.  One run has ~190 test users, i.e. a standard error of ~0.015 on its MRR, so
one seed cannot say whether the oracle clears a floor; this script runs every recipe over --seeds model seeds (the
split stays the reference's) and writes mean, s.d., s.e. and the per-seed values.

    python profiles/tools/oracle_mrr_seeds.py --seeds 16 --out profiles/r2_oracle_mrr_floors.json
"""
import argparse
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))

# (name, model, loss, threads, parallelism, floor default, floor under MKL_CBWR=AVX, citation)
RECIPES = [
    ("lstm_hinge_1thread", "lstm", "hinge", 1, "asynchronous", 0.081, 0.091, "lstm.rs:450-472"),
    ("lstm_hinge_2threads", "lstm", "hinge", 2, "asynchronous", 0.074, 0.078, "lstm.rs:474-496"),
    ("lstm_warp_1thread", "lstm", "warp", 1, "asynchronous", 0.10, 0.089, "lstm.rs:498-520"),
    ("ewma_hinge_1thread", "ewma", "hinge", 1, "asynchronous", 0.11, 0.091, "ewma.rs:463-484"),
    ("ewma_warp_1thread", "ewma", "warp", 1, "asynchronous", 0.14, 0.089, "ewma.rs:486-507"),
]


def split(split_seed=bytes([42] * 16)):
    import oracle_lib as O
    z = np.load(os.path.join(ROOT, "tests", "golden", "ml100k_csr.npz"))
    up = z["user_ptr"].astype(np.int64)
    users = np.repeat(np.arange(944), np.diff(up)).astype(np.uint64)
    items, ts = z["item_ids"].astype(np.uint64), z["timestamps"].astype(np.uint64)
    is_train, _ = O.user_based_split(users, split_seed, 0.2)
    tr = O.compress(users[is_train], items[is_train], ts[is_train], 944)
    te = O.compress(users[~is_train], items[~is_train], ts[~is_train], 944)
    return tr, te


def one(job):
    import oracle_lib as O
    (name, model, loss, threads, par, _, _, _), seed, T, epochs, merge, vary = job
    O.lib().sbo_set_merge_sparse(int(merge))
    tr, te = split(bytes([(seed * 13 + i * 7 + 5) & 255 for i in range(16)]) if vary else bytes([42] * 16))
    m = O.OracleModel(model, 1683, T, embedding_dim=32, learning_rate=0.16, l2_penalty=0.0004, lstm_variant="normal",
                      loss=loss, optimizer="adagrad", parallelism=par, num_threads=threads, num_epochs=epochs,
                      seed=bytes([(seed * 37 + i * 11 + 1) & 255 for i in range(16)]))
    rc, _ = m.fit(tr[0], tr[1])
    assert rc == 0
    rc, mrr = m.mrr_score(te[0], te[1])
    assert rc == 0
    return name, seed, float(mrr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=16)
    ap.add_argument("--seq", type=int, default=128)
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    ap.add_argument("--merge", action="store_true", help="experiment: merged duplicate rows (oracle/sbr_oracle.c g_merge_sparse)")
    ap.add_argument("--vary-split", action="store_true", help="also draw a different user_based_split per seed: the reference's "
                    "floors are single runs on ONE split whose SipHash keys (rand 0.5 Uniform<u64>) we cannot reproduce")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_oracle_mrr_floors.json"))
    a = ap.parse_args()
    import oracle_lib as O
    O.lib()
    recipes = [r for r in RECIPES if not a.only or r[0] in a.only.split(",")]
    jobs = [(r, s, a.seq, a.epochs, a.merge, a.vary_split) for r in recipes for s in range(a.seeds)]
    with mp.Pool(a.procs) as pool:
        res = pool.map(one, jobs, chunksize=1)
    out = {"recipe": "ML-100K, user_based_split(0.2) seed [42;16], seq %d, dim 32, lr 0.16, l2 4e-4, Normal, Adagrad, %d epochs"
                     % (a.seq, a.epochs), "seeds": a.seeds, "split": "a different user_based_split(0.2) per seed" if a.vary_split else "seed [42;16] (lstm.rs:428-430)",
           "sparse_duplicates": "merged per step (experiment)" if a.merge else "un-merged, one visit per recorded entry (default)", "results": {}}
    for r in recipes:
        v = np.array([m for n, _, m in res if n == r[0]])
        se = float(v.std(ddof=1) / np.sqrt(len(v))) if len(v) > 1 else None
        out["results"][r[0]] = {"cite": r[7], "floor_default": r[5], "floor_avx": r[6], "mean": float(v.mean()),
                                "sd": float(v.std(ddof=1)) if len(v) > 1 else None, "se": se,
                                "min": float(v.min()), "max": float(v.max()),
                                "seeds_above_default_floor": int((v > r[5]).sum()),
                                "seeds_above_avx_floor": int((v > r[6]).sum()), "values": [float(x) for x in v]}
        print(r[0], "mean %.4f sd %.4f se %.4f  floors %.3f / %.3f  above: %d / %d of %d" % (
            v.mean(), v.std(ddof=1) if len(v) > 1 else 0, se or 0, r[5], r[6], (v > r[5]).sum(), (v > r[6]).sum(), len(v)))
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
