"""Stand-alone embedding-gather bandwidth (BASELINE.json: "embed-gather HBM GB/s vs roofline"): gather_rows_kernel on
HBM-resident tables, uniform random ids, device-timed (sbr_model_gather_rows_timed).  Algorithmic bytes per row:
4 D read + 4 D written + 4 (u32 id).  usage (GPU box): python profiles/tools/gather_bw.py"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
PEAK = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
for N, D, rows in [(1_000_000, 64, 1 << 23), (8_000_000, 128, 1 << 22)]:
    m = pkg.ewma.Hyperparameters(N, 8).embedding_dim(D).optimizer(pkg.Optimizer.Adagrad).num_threads(1).from_seed(bytes(range(16))).build()
    ids = np.random.default_rng(1).integers(0, N, size=rows, dtype=np.uint64)
    ms = m.gather_rows_timed(ids, iters=5)
    gbs = rows * (8 * D + 4) / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "gather_rows_kernel", "num_items": N, "dim": D, "rows": rows, "table_GB": N * 2 * D * 4 / 1e9,
                      "kernel_ms": round(ms, 4), "alg_GBps": round(gbs, 1), "frac_of_measured_hbm_peak": round(gbs / PEAK, 3)}), flush=True)
    del m
