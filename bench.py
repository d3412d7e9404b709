#!/usr/bin/env python
"""bench.py -- user-sequence training steps/s of the fit() hot path on B200 (BASELINE.json metric).

A "step" of this benchmark is one pass (one epoch of fit) over a synthetic ML-100K-shaped interaction stream
(configs[1] of BASELINE.json: 1,683 items, dim 32, every user-sequence 32 items => 31 timesteps, LSTM Normal,
WARP, Adagrad lr 0.16 l2 4e-4).  `value` counts reference optimizer steps (= sub-sequences, sequence_model.rs:111)
per second with the stream already resident in HBM; `e2e` is the same metric through the public C ABI from HOST
buffers (CSR -> sbr_compressed_from_csr -> sbr_model_fit: H2D of the id stream and the schedule inside the timed
region).  `--impl reference` times the CPU restatement of the reference path (oracle/, all host threads).

Contract: python bench.py --gpus N --steps K --warmup W ; N > 1 under torchrun (one rank per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NUM_ITEMS, SEQ_LEN, DIM = 1683, 32, 32
LR, L2 = 0.16, 4e-4
A_TRAIN_BYTES_PER_TIMESTEP = 60 * DIM + 52  # gather 12D+20 + Adagrad visit 48D+32  (= 1972 B at D=32)
METRIC = "user-seq steps/sec"


def make_stream(num_seqs, seed, zipf=False):
    """2^k synthetic users x exactly 32 items (SURVEY 8d 'C2-stream'): uniform over [1, N) -- the headline variant -- or
    Zipf(1.0) over the same ids (p(rank k) ~ 1/k: the hot-row case, item 1 alone draws 13 % of all interactions)."""
    rng = np.random.default_rng(seed)
    ptr = np.arange(num_seqs + 1, dtype=np.uint64) * np.uint64(SEQ_LEN)
    if zipf:
        cdf = np.cumsum(1.0 / np.arange(1, NUM_ITEMS, dtype=np.float64))
        ids = (1 + np.searchsorted(cdf / cdf[-1], rng.random(num_seqs * SEQ_LEN))).astype(np.uint64)
        ids = np.minimum(ids, NUM_ITEMS - 1)
    else:
        ids = rng.integers(1, NUM_ITEMS, size=num_seqs * SEQ_LEN, dtype=np.uint64)
    return ptr, ids


def gather_leg(pkg):
    """BASELINE.json metric, second leg: "embed-gather HBM GB/s vs roofline".  Stand-alone gather_rows_kernel on an
    HBM-resident table (8 M items x dim 128 with its Adagrad state: 8 GB, far beyond L2), 4 M uniform random row ids per
    launch, timed by the library with CUDA events on its stream (sbr_model_gather_rows_timed, mean of 5 launches after a
    warm-up).  Algorithmic bytes per row: 4 D read + 4 D written + 4 (u32 id)."""
    N, D, rows = 8 << 20, 128, 4 << 20
    m = (pkg.ewma.Hyperparameters(N, 8).embedding_dim(D).optimizer(pkg.Optimizer.Adagrad).from_seed(bytes(range(16)))).build()
    ids = np.random.default_rng(7).integers(0, N, size=rows, dtype=np.uint64)
    ms = m.gather_rows_timed(ids, iters=5)
    nbytes = rows * (8 * D + 4)
    del m
    return {"GB/s": nbytes / (ms * 1e-3) / 1e9, "ms_per_launch": ms, "rows_per_launch": rows, "num_items": N, "dim": D,
            "table_bytes": N * D * 4 * 2, "algorithmic_bytes_per_row": 8 * D + 4, "kernel": "gather_rows_kernel"}


def sharded_c4_leg(pkg, torch, dist, rank, world, steps, warmup, barrier, max_over_ranks):
    """BASELINE configs[3] / north_star's multi-GPU requirement: EWMA dim=128 seq=128 BPR Adagrad on a catalogue that is
    ROW-SHARDED over the GPUs (item id % world; 6.25 M items = 6.5 GB of records per GPU, 50 M items at 8 GPUs: weak scaling),
    trained by the round-synchronous engine (Parallelism::Synchronous): per round the requested (row, order) pairs, the rows
    and the gradient rows cross NVLink as NCCL all-to-alls (grouped ncclSend/ncclRecv), optimizer state never moves.  Every
    rank trains its own users; `value` = sub-sequences of all ranks / max-over-ranks wall time, device-resident plan."""
    items_per_gpu, D4, T4, P4, rounds = 6_250_000, 128, 128, 4736, 8
    N4 = items_per_gpu * world
    S4 = P4 * rounds
    rng = np.random.default_rng(4000 + rank)
    ptr = np.arange(S4 + 1, dtype=np.uint64) * np.uint64(T4)
    ids = rng.integers(1, N4, size=S4 * T4, dtype=np.uint64)
    h = (pkg.ewma.Hyperparameters(N4, T4).embedding_dim(D4).learning_rate(0.05).l2_penalty(0.0).loss(pkg.Loss.BPR)
         .optimizer(pkg.Optimizer.Adagrad).parallelism(pkg.Parallelism.Synchronous).num_epochs(1).num_threads(P4)
         .from_seed(bytes(range(16))))
    if world > 1:
        h = h.shard(rank, world)
    model = h.build()
    plan = model.fit_plan(pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N4).upload())
    for _ in range(max(1, warmup - 1)):
        plan.run()
    barrier()
    t0 = time.perf_counter()
    kms, ts = 0.0, 0
    for _ in range(steps):
        plan.run()
        st = plan.stats()
        kms += st["train_kernel_ms"]; ts += st["timesteps"]
    barrier()
    wall = max_over_ranks(time.perf_counter() - t0)
    st = plan.stats()
    a_bytes = 60 * D4 + 52
    peak = peaks()[0]
    out = {"workload": "EWMA dim=128 seq=128 BPR Adagrad, item table row-sharded over %d GPU(s) (id %% world), %d items (BASELINE configs[3], weak-scaled: "
                       "%d items per GPU)" % (world, N4, items_per_gpu),
           "value": world * st["steps"] * steps / wall, "unit": "steps/s", "ms_per_step": wall / steps * 1e3,
           "seqs_per_gpu_per_step": int(st["steps"]), "partitions_per_gpu": int(st["partitions"]), "rounds_per_step": rounds,
           "kernel": st["kernel"], "gpu_launches_per_step": int(st["kernel_launches"]),
           "per_gpu_hbm_frac": a_bytes * ts / (kms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_timestep": a_bytes,
           "exchange": ("none (one GPU owns every row)" if world == 1 else
                        "(row, order) pairs: NCCL grouped send/recv all-to-all; rows + biases and gradient rows: pushed into the peers' buffers over NVLink "
                        "(CUDA IPC mappings) by the copy engines, one stream per peer, one 4-byte all-reduce per phase as the barrier; two half-round "
                        "pipelines on two streams" if "p2p copy" in st["kernel"] else
                        "(row, order) pairs: NCCL grouped send/recv all-to-all; rows + biases and gradient rows: plain stores of the owners' gather kernel / the "
                        "requesters' compute kernel into peer buffers over NVLink (CUDA IPC mappings), one 4-byte all-reduce per phase as the barrier; two "
                        "half-round pipelines on two streams" if "p2p" in st["kernel"] else
                        "NCCL grouped send/recv all-to-all of (row, order) pairs, rows + biases, gradient rows; two half-round pipelines on two streams")}
    del plan, model
    return out


def wide_lstm_leg(pkg, name, rank, world, steps, barrier, max_over_ranks):
    """BASELINE configs[2] (C3: 1 M items, LSTM dim 64 seq 64 Hinge Adam) and configs[4] (C5: ML-20M-shaped, 27 K items, LSTM dim 256
    seq 200 WARP Adagrad): the batched round engine on tcgen05 GEMMs (lstm_batch.cuh), device-resident plan, one epoch over a
    synthetic stream per step.  N > 1: a full replica per GPU on its own users, sbr_model_replica_sync (one ncclAllReduce of the
    parameter / optimizer-state deltas) after every step, inside the timed region -- the "1 vs 8 scaling" of configs[4]."""
    N, D, L, S, loss, opt, lr = {"c3": (1_000_000, 64, 64, 1 << 19, pkg.Loss.Hinge, pkg.Optimizer.Adam, 0.01),
                                 "c5": (27_000, 256, 200, 1 << 17, pkg.Loss.WARP, pkg.Optimizer.Adagrad, 0.16)}[name]
    rng = np.random.default_rng(5000 + rank)
    ptr = np.arange(S + 1, dtype=np.uint64) * np.uint64(L)
    ids = rng.integers(1, N, size=S * L, dtype=np.uint64)
    model = (pkg.lstm.Hyperparameters(N, L).embedding_dim(D).learning_rate(lr).l2_penalty(4e-4).loss(loss).optimizer(opt)
             .lstm_variant(pkg.LSTMVariant.Normal).parallelism(pkg.Parallelism.Asynchronous).num_epochs(1).num_threads(0)
             .from_seed(bytes(range(16))).build())
    plan = model.fit_plan(pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N).upload())
    if world > 1:
        model.replica_sync()
    plan.run()
    if world > 1:
        model.replica_sync()
    barrier()
    t0 = time.perf_counter()
    kms, ts = 0.0, 0
    for _ in range(steps):
        plan.run()
        if world > 1:
            model.replica_sync()
        st = plan.stats()
        kms += st["train_kernel_ms"]; ts += st["timesteps"]
    barrier()
    wall = max_over_ranks(time.perf_counter() - t0)
    st = plan.stats()
    a_bytes = (84 * D + 68) if name == "c3" else (60 * D + 52)
    out = {"value": world * st["steps"] * steps / wall, "unit": "steps/s", "ms_per_step": wall / steps * 1e3,
           "timesteps_per_s": world * ts / wall, "gemm_tflops_algorithmic": world * ts / wall * 48 * D * D / 1e12,
           "seqs_per_gpu_per_step": int(st["steps"]), "partitions_per_gpu": int(st["partitions"]), "kernel": st["kernel"],
           "gpu_launches_per_step": int(st["kernel_launches"]), "per_gpu_hbm_frac": a_bytes * ts / (kms * 1e-3) / 1e9 / peaks()[0],
           "algorithmic_bytes_per_timestep": a_bytes,
           "workload": {"c3": "synthetic 1M items, LSTM Normal dim=64 seq=64 Hinge Adam (BASELINE configs[2])",
                        "c5": "ML-20M-shaped 27K items, LSTM Normal dim=256 seq=200 WARP Adagrad (BASELINE configs[4])"}[name]}
    del plan, model
    return out


def learnable_task_leg(pkg):
    """Convergence evidence at the benchmarked concurrency (the uniform-random bench stream has nothing to learn): a catalogue of
    the same shape whose sequences follow a noisy item -> item map (next = perm[cur] with p = 0.8, else uniform).  Test MRR of
    the next item for 4,096 held-out users (mrr_score) after the automatic schedule's bounded first epoch and after two more
    device-filling epochs, beside a model started cold at the device-filling partition count."""
    N, L, S = NUM_ITEMS, SEQ_LEN, 1 << 20
    rng = np.random.default_rng(77)
    perm = rng.permutation(np.arange(1, N))

    def chains(users, seed):
        r = np.random.default_rng(seed)
        out = np.empty((users, L), dtype=np.uint64)
        cur = r.integers(1, N, size=users)
        for t in range(L):
            out[:, t] = cur
            cur = np.where(r.random(users) >= 0.8, r.integers(1, N, size=users), perm[cur - 1])
        return np.arange(users + 1, dtype=np.uint64) * np.uint64(L), out.reshape(-1)

    ptr, ids = chains(S, 1)
    tptr, tids = chains(4096, 2)
    train = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=N).upload()
    test = pkg.CompressedInteractions.from_csr(tptr, tids, None, num_items=N).upload()

    def build(threads):
        return (pkg.lstm.Hyperparameters(N, L).embedding_dim(DIM).learning_rate(0.05).l2_penalty(0.0).loss(pkg.Loss.WARP)
                .optimizer(pkg.Optimizer.Adagrad).lstm_variant(pkg.LSTMVariant.Normal).parallelism(pkg.Parallelism.Asynchronous)
                .num_epochs(1).num_threads(threads).from_seed(bytes(range(16))).build())
    m = build(0)
    hist = []
    for _ in range(3):
        m.fit(train)
        hist.append({"partitions": int(m.last_fit_stats()["partitions"]), "test_mrr": float(pkg.mrr_score(m, test))})
    device_fill = hist[-1]["partitions"]
    cold = build(device_fill)
    for _ in range(3):
        cold.fit(train)
    return {"task": "noisy item->item chain, %d items, 2^20 x %d training users, 4096 test users; LSTM dim %d WARP Adagrad lr 0.05" % (N, L, DIM),
            "automatic_num_threads_epochs": hist, "cold_start_at_%d_partitions_after_3_epochs_test_mrr" % device_fill: float(pkg.mrr_score(cold, test)),
            "untrained_mrr": 0.005, "ceiling_mrr": 0.79}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  nvidia-smi needs
    ~100 ms to start, longer than a short timed region: the sampler is started before the warm-up (20 ms period), every row
    is stamped on arrival, and only rows that arrived inside [begin(), end()] are reported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.03)   # a row in flight
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for ts, r in self.rows if self.t0 is not None and self.t0 <= ts <= (self.t1 or ts) + 0.03]
        in_region = len(inside)
        if not inside and self.rows and self.t0 is not None:   # a region shorter than the sampling period: the rows closest to it
            mid = 0.5 * (self.t0 + (self.t1 or self.t0))
            inside = [r for _, r in sorted(self.rows, key=lambda tr: abs(tr[0] - mid))[:2]]
        for r in inside:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for k, n in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm), "samples_inside_timed_region": in_region}


def cpu_reference_run(num_seqs, steps, warmup, threads, seed=1234):
    """The reference arm / cpu_baseline: the oracle's fit() (sequence_model.rs:70-178 restated in C) with all host
    threads, lock-free Hogwild (Parallelism::Asynchronous, the fastest reference mode), on a bounded sample of the
    same workload.  Timed like the reference's tests time fit() (lstm.rs:436-438)."""
    import oracle_lib as O
    ptr, ids = make_stream(num_seqs, seed)
    m = O.OracleModel("lstm", NUM_ITEMS, SEQ_LEN, embedding_dim=DIM, learning_rate=LR, l2_penalty=L2,
                      lstm_variant="normal", loss="warp", optimizer="adagrad", parallelism="asynchronous",
                      num_threads=threads, num_epochs=1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        rc, _ = m.fit(ptr, ids)
        assert rc == 0
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return num_seqs * len(times) / sum(times), sum(times) / len(times)


class ReplicaSync:
    """N > 1 on a small (L2-resident) catalogue: every GPU trains a FULL replica on its own users (local Hogwild, the
    single-GPU kernel at full speed) and after every step the replicas exchange what they changed: the deltas of all
    parameters and of their Adagrad accumulators are summed over the ranks with one NCCL all-reduce and applied to the
    common starting point, so that every rank continues from the same model -- the sum of all partitions' updates, as
    in one big Hogwild run, with a staleness of one step instead of ~0.  (Sharing ONE row-sharded model through NVLink
    peer mappings -- `SBR_BENCH_MULTI=shared`, DESIGN.md 3.6 -- makes every second row access a small remote request:
    2 GPUs then run slower than one.)  Host plumbing only: get/set_parameter of the public API + torch.distributed."""

    def __init__(self, model, names, torch, dist):
        self.m, self.names, self.torch, self.dist = model, names, torch, dist
        self.dev = "cuda" if torch.cuda.is_available() else "cpu"   # (cpu + gloo in tests/dist_worker.py)
        self.prev = [model.get_parameter(n) for n in names]
        self.sizes = [len(p) for p in self.prev]
        self.bytes_per_sync = 4 * sum(self.sizes)

    def __call__(self):
        cur = [self.m.get_parameter(n) for n in self.names]
        delta = self.torch.from_numpy(np.concatenate([c - p for c, p in zip(cur, self.prev)])).to(self.dev)
        self.dist.all_reduce(delta)
        delta = delta.cpu().numpy()
        off = 0
        for i, n in enumerate(self.names):
            new = self.prev[i] + delta[off:off + self.sizes[i]]
            off += self.sizes[i]
            self.m.set_parameter(n, new)
            self.prev[i] = new


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seqs", type=int, default=1 << 20, help="sub-sequences per GPU per step")
    ap.add_argument("--cpu-seqs", type=int, default=0, help="sample size of the CPU arm (0 = auto)")
    ap.add_argument("--threads", type=int, default=0, help="Hogwild partitions per GPU (0 = fill the device)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    config = {"workload": "ML-100K-shaped stream: LSTMVariant::Normal dim=32 seq=32 WARP Adagrad lr=0.16 l2=4e-4 "
                          "(BASELINE configs[1])", "num_items": NUM_ITEMS, "seq_len": SEQ_LEN, "dim": DIM,
              "item_distribution": "uniform[1,N)"}

    if args.impl == "reference":
        if rank != 0:
            return
        n = args.cpu_seqs or 2048 * cores
        value, sec = cpu_reference_run(n, args.steps, args.warmup, cores)
        config.update({"seqs_per_step": n, "parallelism": "%d host threads, Hogwild" % cores})
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
                             "sample": "%d sequences x %d items per step, oracle fit(), %d threads Hogwild" % (n, SEQ_LEN, cores)},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    pkg = g.load_package()
    if pkg.device_count() < 1:
        raise SystemExit("bench.py needs a B200: libsbr_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    pkg.set_device(local_rank)
    sampler = ClockSampler(local_rank)   # (nvidia-smi takes a few hundred ms to deliver its first row: started long before the timed region)
    sampler.start()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    S = args.seqs
    ptr, ids = make_stream(S, 1000 + rank)  # every rank trains on its own shard of the stream (weak scaling)
    # N > 1: ONE model shared by all ranks -- the item table is row-sharded (id % N) and every rank's kernel reads /
    # updates remote rows in the owner's HBM over NVLink (CUDA-IPC peer mappings); no collective on the data path.
    seed = bytes(range(16))  # same init on every rank
    def hyper_factory():
        return (pkg.lstm.Hyperparameters(NUM_ITEMS, SEQ_LEN).embedding_dim(DIM).learning_rate(LR).l2_penalty(L2)
                .lstm_variant(pkg.LSTMVariant.Normal).loss(pkg.Loss.WARP).optimizer(pkg.Optimizer.Adagrad)
                .parallelism(pkg.Parallelism.Asynchronous).num_epochs(1).num_threads(args.threads).from_seed(seed))
    hyper = hyper_factory()
    multi = os.environ.get("SBR_BENCH_MULTI", "replicas") if world > 1 else "single"
    sync = None
    if world > 1:   # the library's own NCCL communicator (sbr_dist_init): replica all-reduce and the sharded engine's all-to-alls
        uid = [pkg.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        pkg.dist_init(rank, world, uid[0])
    if multi == "shared":
        model = hyper.shard(rank, world).build()
        blobs = [None] * world
        dist.all_gather_object(blobs, model.ipc_export())
        model.ipc_attach(blobs)
        dist.barrier()
    else:
        model = hyper.build()
        if world > 1:
            class LibSync:   # sbr_model_replica_sync: one ncclAllReduce on device buffers inside the library
                bytes_per_sync = 0
                def __call__(self):
                    self.bytes_per_sync = model.replica_sync() or self.bytes_per_sync
            sync = LibSync()
            sync()           # records the common starting point

    # ---------------- device-resident arm: `value` ----------------
    data = pkg.CompressedInteractions.from_csr(ptr, ids, None, num_items=NUM_ITEMS).upload()
    # Warm start (untimed): with num_threads = 0 a cold LSTM trains its first epoch at a bounded partition count
    # (2.5 per item: 4,096 here) and fills the device afterwards -- from random parameters 37,888 concurrent sequences on
    # 1,683 item rows do not learn.  The measured steps are device-filling epochs of a model that has had that first epoch.
    model.fit(data)
    warm_partitions = model.last_fit_stats()["partitions"]
    plan = model.fit_plan(data)
    for _ in range(args.warmup):
        plan.run()
        if sync:
            sync()
    barrier()
    sampler.begin()
    t0 = time.perf_counter()
    kernel_ms, launches, timesteps = 0.0, 0, 0
    for _ in range(args.steps):
        plan.run()  # blocks until the epoch's kernel has finished (loss read back)
        if sync:
            sync()
        st = plan.stats()
        kernel_ms += st["train_kernel_ms"]; launches += st["kernel_launches"]; timesteps += st["timesteps"]
    barrier()
    sampler.end()
    wall = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop()
    partitions = plan.stats()["partitions"]
    steps_done = plan.stats()["steps"]  # per run
    kernel_name = plan.stats()["kernel"]
    value = world * steps_done * args.steps / wall
    kernel_ms_max = max_over_ranks(kernel_ms)
    del plan

    # ---------------- Zipf(1.0) variant of the same stream (SURVEY 8d "report both"), device-resident, rank 0's number ----
    zipf = None
    if world == 1:
        zptr, zids = make_stream(S, 2000 + rank, zipf=True)
        zdata = pkg.CompressedInteractions.from_csr(zptr, zids, None, num_items=NUM_ITEMS).upload()
        zmodel = hyper_factory().build()
        zmodel.fit(zdata)
        zplan = zmodel.fit_plan(zdata)
        for _ in range(args.warmup):
            zplan.run()
        zms, zts = 0.0, 0
        for _ in range(args.steps):
            zplan.run()
            zms += zplan.stats()["train_kernel_ms"]; zts += zplan.stats()["timesteps"]
        zipf = {"value": zplan.stats()["steps"] * args.steps / (zms * 1e-3), "unit": "steps/s", "timed": "CUDA events around the kernel",
                "frac": A_TRAIN_BYTES_PER_TIMESTEP * zts / (zms * 1e-3) / 1e9 / peaks()[0], "item_distribution": "Zipf(1.0) over [1,N)"}
        del zplan, zmodel, zdata, zids

    # ---------------- end-to-end arm through the C ABI from host buffers: `e2e` ----------------
    # The CSR (usize user_pointers / item_ids, as the reference holds them) lives in page-locked host memory; every step
    # hands it to the library, which moves the id stream to HBM, builds the schedule (device chunker + host master shuffle),
    # trains one epoch and returns the loss.  The same from ordinary pageable numpy arrays is reported as `pageable_value`.
    def e2e_arm(ptr_h, ids_h):
        h2d_ = d2h_ = 0
        up_ms = prep_ms = 0.0
        for _ in range(2):
            model.fit(pkg.CompressedInteractions.from_csr(ptr_h, ids_h, None, num_items=NUM_ITEMS, borrow=True))
            if sync:
                sync()
        barrier()
        t0_ = time.perf_counter()
        for _ in range(args.steps):
            c = pkg.CompressedInteractions.from_csr(ptr_h, ids_h, None, num_items=NUM_ITEMS, borrow=True)  # host CSR in, nothing resident
            model.fit(c)
            if sync:
                sync()
            st_ = model.last_fit_stats()
            h2d_, d2h_ = st_["h2d_bytes"], st_["d2h_bytes"]
            up_ms += st_["upload_ms"]; prep_ms += st_["host_prepare_ms"]
            del c
        barrier()
        wall_ = max_over_ranks(time.perf_counter() - t0_)
        return world * steps_done * args.steps / wall_, h2d_, d2h_, up_ms / args.steps, prep_ms / args.steps

    pin_ids = torch.empty(len(ids), dtype=torch.int64).pin_memory()
    pin_ptr = torch.empty(len(ptr), dtype=torch.int64).pin_memory()
    ids_pl, ptr_pl = pin_ids.numpy().view(np.uint64), pin_ptr.numpy().view(np.uint64)
    ids_pl[:] = ids; ptr_pl[:] = ptr
    e2e_value, h2d, d2h, e2e_upload_ms, e2e_host_ms = e2e_arm(ptr_pl, ids_pl)
    e2e_pageable = e2e_arm(ptr, ids)

    del model
    c4 = None
    if os.environ.get("SBR_BENCH_C4", "1") != "0":
        c4 = sharded_c4_leg(pkg, torch, dist, rank, world, args.steps, args.warmup, barrier, max_over_ranks)

    learn = learnable_task_leg(pkg) if (rank == 0 and os.environ.get("SBR_BENCH_LEARN", "1") != "0") else None
    wide = {}
    if os.environ.get("SBR_BENCH_WIDE", "1") != "0":
        for cname in ("c3", "c5"):
            wide[cname] = wide_lstm_leg(pkg, cname, rank, world, max(2, min(args.steps, 3)), barrier, max_over_ranks)

    if rank == 0:
        peak, peak_kind = peaks()
        per_launch_bytes = A_TRAIN_BYTES_PER_TIMESTEP * (timesteps / max(launches, 1))
        # DRAM bytes per launch: NOT measured in this run -- taken from the committed `ncu --set full` capture of the same
        # kernel on the same workload (dram__bytes_read.sum + dram__bytes_write.sum), scaled by sequences per launch
        traffic, traffic_src = None, None
        for tname in ("r2_traffic.json", "r1_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath):
                with open(tpath) as f:
                    tj = json.load(f)
                traffic = tj["dram_bytes_per_launch"] * (steps_done / tj["seqs_per_launch"])
                traffic_src = "profiles/%s (ncu --set full capture of %s, %d sequences per launch; scaled to this launch)" % (
                    tname, tj.get("kernel", "the tile kernel"), tj["seqs_per_launch"])
                break
        per_launch_ms = kernel_ms / max(launches, 1)
        achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config, seqs_per_gpu_per_step=S, partitions_per_gpu=int(partitions),
                           warm_start="one untimed epoch at %d partitions (automatic num_threads on a cold LSTM, DESIGN 4.5), then device-filling" % warm_partitions,
                           parallelism=("hogwild partitions; one shared model, item table row-sharded over %d GPUs via NVLink peer access" % world)
                           if multi == "shared" else
                           ("hogwild partitions per GPU; full replica per GPU, deltas of parameters and Adagrad state summed over %d GPUs "
                            "(one ncclAllReduce of %d bytes on device buffers inside the library) after every step" % (world, sync.bytes_per_sync)) if world > 1 else "hogwild partitions",
                           l2_policy="id stream (%d MiB/GPU) larger than L2; 215 KB item table is L2-resident by construction"
                                     % (S * SEQ_LEN * 4 >> 20)),
            "timesteps_per_s": value * (SEQ_LEN - 1),
            "device_ms_per_step": kernel_ms_max / args.steps,
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "host_buffers": "page-locked usize CSR (raw 64-bit ids DMA'ed, narrowed on the device)",
                    "id_upload_ms": e2e_upload_ms, "host_schedule_ms": e2e_host_ms,
                    "pageable_value": e2e_pageable[0], "pageable_h2d_bytes_per_step": int(e2e_pageable[1])},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind, "kernel": kernel_name,
                         "algorithmic_bytes_per_timestep": A_TRAIN_BYTES_PER_TIMESTEP},
        }
        if c4:
            out["sharded_c4"] = c4
        for cname, leg in wide.items():
            out[cname] = leg
        if learn:
            out["convergence"] = learn
        if zipf:
            out["zipf"] = zipf
        if world == 1:
            gl = gather_leg(pkg)
            gl["peak"] = peak; gl["frac"] = gl["GB/s"] / peak
            out["gather"] = gl
        if not args.no_cpu_baseline:
            n = args.cpu_seqs or 1024 * cores
            cv, csec = cpu_reference_run(n, 2, 1, cores)
            out["cpu_baseline"] = {"value": cv, "unit": "steps/s", "cores": cores, "kind": "port",
                                   "sample": "%d sequences x %d items, oracle fit() 1 epoch x2 (+1 warm-up), %d threads Hogwild"
                                             % (n, SEQ_LEN, cores)}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        pkg.dist_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
