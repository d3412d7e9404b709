"""steps/s of the LSTM tile kernels as a function of the item-table size (hot-line / TLB experiments)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
S, L, D = 1 << 20, 32, 32
def run(N, gen, loss="warp", reps=3, dbg="0"):
    os.environ["SBR_DBG_FLAGS"] = dbg
    os.environ["SBR_LSTM_TC"] = gen
    rng = np.random.default_rng(1)
    ptr = np.arange(S + 1, dtype=np.uint64) * np.uint64(L)
    ids = rng.integers(1, N, size=S * L, dtype=np.uint64)
    h = (pkg.lstm.Hyperparameters(N, L).embedding_dim(D).learning_rate(0.16).l2_penalty(4e-4)
         .lstm_variant(pkg.LSTMVariant.Normal).loss({"warp": pkg.Loss.WARP, "hinge": pkg.Loss.Hinge, "bpr": pkg.Loss.BPR}[loss])
         .optimizer(pkg.Optimizer.Adagrad).parallelism(pkg.Parallelism.Asynchronous).num_epochs(1).num_threads(0).from_seed(bytes(range(16))))
    model = h.build()
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N).upload()
    plan = model.fit_plan(data)
    for _ in range(2): plan.run()
    ms = []
    for _ in range(reps):
        plan.run(); ms.append(plan.stats()["train_kernel_ms"])
    st = plan.stats()
    print("dbg=%s N=%8d gen=%-2s loss=%s kernel_ms=%.2f steps/s=%.3e" % (dbg, N, gen, loss, min(ms), st["steps"] / (min(ms) * 1e-3)), flush=True)
losses = sys.argv[3].split(",") if len(sys.argv) > 3 else ["warp"]
dbgs = sys.argv[4].split(",") if len(sys.argv) > 4 else ["0"]
for gen in sys.argv[1].split(","):
    for N in [int(x) for x in sys.argv[2].split(",")]:
        for loss in losses:
            for dbg in dbgs:
                run(N, gen, loss, dbg=dbg)
