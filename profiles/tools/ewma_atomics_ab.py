"""A/B of the EWMA kernel with L2-atomic vs plain read-modify-write Adagrad visits (SBR_DBG_FLAGS=8) at the C1-stream
and C4-slice shapes.  usage (GPU box): python profiles/tools/ewma_atomics_ab.py"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
def run(N, D, L, S, loss, dbg, threads=0):
    os.environ["SBR_DBG_FLAGS"] = dbg
    rng = np.random.default_rng(1)
    ptr = np.arange(S + 1, dtype=np.uint64) * np.uint64(L)
    ids = rng.integers(1, N, size=S * L, dtype=np.uint64)
    h = (pkg.ewma.Hyperparameters(N, L).embedding_dim(D).learning_rate(0.16).l2_penalty(4e-4)
         .loss({"warp": pkg.Loss.WARP, "hinge": pkg.Loss.Hinge, "bpr": pkg.Loss.BPR}[loss])
         .optimizer(pkg.Optimizer.Adagrad).parallelism(pkg.Parallelism.Asynchronous).num_epochs(1).num_threads(threads).from_seed(bytes(range(16))))
    model = h.build()
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N).upload()
    plan = model.fit_plan(data)
    for _ in range(2): plan.run()
    ms = []
    for _ in range(3):
        plan.run(); ms.append(plan.stats()["train_kernel_ms"])
    st = plan.stats()
    A = 60 * D + 52
    print("EWMA dbg=%s N=%d D=%d L=%d S=%d loss=%s P=%d kernel_ms=%.2f steps/s=%.3e algGB/s=%.0f" % (dbg, N, D, L, S, loss, st["partitions"], min(ms), st["steps"] / (min(ms) * 1e-3), A * st["timesteps"] / (min(ms) * 1e-3) / 1e9), flush=True)
for dbg in ("0", "8"):
    run(1683, 32, 32, 1 << 20, "bpr", dbg)
    run(1683, 32, 32, 1 << 20, "warp", dbg)
    run(50_000_000, 128, 128, 1 << 16, "bpr", dbg)
