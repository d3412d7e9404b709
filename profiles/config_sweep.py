"""Kernel-time throughput of every BASELINE.json config shape on ONE B200 through the C ABI (device-resident plan).
Prints one line per config: steps/s, timesteps/s and the algorithmic HBM fraction (SURVEY 8d bytes / measured peak).
usage: python profiles/config_sweep.py [c1 c2 c3 c4 c5]      (results quoted in BASELINE.md section 4)"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
CONFIGS = {  # kind, N, D, L, sequences, loss, optimizer, zipf
    "c1": ("ewma", 1683, 32, 32, 1 << 20, "bpr", "adagrad"),
    "c2": ("lstm", 1683, 32, 32, 1 << 20, "warp", "adagrad"),
    "c3": ("lstm", 1_000_000, 64, 64, 1 << 19, "hinge", "adam"),
    "c4": ("ewma", 50_000_000, 128, 128, 1 << 16, "bpr", "adagrad"),
    "c5": ("lstm", 27_000, 256, 200, 1 << 17, "warp", "adagrad"),
}
LOSS = {"warp": pkg.Loss.WARP, "hinge": pkg.Loss.Hinge, "bpr": pkg.Loss.BPR}
OPT = {"adagrad": pkg.Optimizer.Adagrad, "adam": pkg.Optimizer.Adam}
for name in (sys.argv[1:] or list(CONFIGS)):
    kind, N, D, L, S, loss, opt = CONFIGS[name]
    S = int(os.environ.get("SBR_SWEEP_SEQS", S))   # (profiling runs use a smaller stream)
    rng = np.random.default_rng(42)
    ptr = np.arange(S + 1, dtype=np.uint64) * np.uint64(L)
    ids = rng.integers(1, N, size=S * L, dtype=np.uint64)
    H = pkg.lstm.Hyperparameters if kind == "lstm" else pkg.ewma.Hyperparameters
    h = (H(N, L).embedding_dim(D).learning_rate(0.16 if opt == "adagrad" else 0.01).l2_penalty(4e-4).loss(LOSS[loss]).optimizer(OPT[opt])
         .parallelism(pkg.Parallelism.Asynchronous).num_epochs(1).num_threads(0).from_seed(bytes(range(16))))
    if kind == "lstm":
        h = h.lstm_variant(pkg.LSTMVariant.Normal)
    model = h.build()
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N).upload()
    if kind == "lstm" and N < 10_000:
        model.fit(data)    # automatic num_threads holds a cold LSTM on a small catalogue to 2.5 partitions per item for its first epoch
    plan = model.fit_plan(data)
    plan.run()
    ms = []
    for _ in range(3):
        plan.run(); ms.append(plan.stats()["train_kernel_ms"])
    st = plan.stats()
    A = (60 * D + 52) if opt == "adagrad" else (84 * D + 68)
    sec = min(ms) * 1e-3
    print(json.dumps({"config": name, "model": kind, "N": N, "D": D, "L": L, "sequences": S, "loss": loss, "optimizer": opt,
                      "partitions": st["partitions"], "kernel_ms": round(min(ms), 3), "steps_per_s": st["steps"] / sec,
                      "timesteps_per_s": st["timesteps"] / sec, "alg_bytes_per_timestep": A,
                      "alg_hbm_frac": A * st["timesteps"] / sec / 1e9 / PEAK}), flush=True)
    del plan, model
