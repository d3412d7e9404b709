"""Where an end-to-end fit() from a host CSR spends its time (host schedule, device span, kernel): prints sbr_fit_stats
per call for the bench workload.  usage (GPU box): python profiles/tools/e2e_breakdown.py"""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
S, L, N, D = 1 << 20, 32, 1683, 32
rng = np.random.default_rng(1)
ptr = np.arange(S + 1, dtype=np.uint64) * np.uint64(L)
ids = rng.integers(1, N, size=S * L, dtype=np.uint64)
h = (pkg.lstm.Hyperparameters(N, L).embedding_dim(D).learning_rate(0.16).l2_penalty(4e-4).lstm_variant(pkg.LSTMVariant.Normal)
     .loss(pkg.Loss.WARP).optimizer(pkg.Optimizer.Adagrad).parallelism(pkg.Parallelism.Asynchronous).num_epochs(1).num_threads(0).from_seed(bytes(range(16))))
model = h.build()
for i in range(5):
    t0 = time.perf_counter()
    c = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N, borrow=True)
    t1 = time.perf_counter()
    model.fit(c)
    t2 = time.perf_counter()
    st = model.last_fit_stats()
    print("from_csr %.1f ms | fit %.1f ms | host_prepare %.1f | total_device %.1f | kernel %.1f | h2d %.1f MB" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, st["host_prepare_ms"], st["total_device_ms"], st["train_kernel_ms"], st["h2d_bytes"] / 1e6), flush=True)
    del c
