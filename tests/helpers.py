"""Shared helpers for the parity tests (GPU engine vs CPU oracle on identical inputs)."""
import numpy as np

LOSSES = {"bpr": 0, "hinge": 1, "warp": 2}
OPTS = {"adagrad": 0, "adam": 1}
VARIANTS = {"normal": 0, "coupled": 1}
PARS = {"asynchronous": 0, "synchronous": 1}


def random_csr(rng, num_users, num_items, min_len, max_len, first_item=1):
    lens = rng.integers(min_len, max_len + 1, size=num_users)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    ids = rng.integers(first_item, num_items, size=int(ptr[-1])).astype(np.uint64)
    return ptr, ids


def stream_csr(rng, num_users, num_items, length, zipf=False):
    """ML-100K-shaped synthetic stream: every user has exactly `length` items."""
    ptr = (np.arange(num_users + 1, dtype=np.uint64) * np.uint64(length))
    n = num_users * length
    if zipf:
        w = 1.0 / np.arange(1, num_items, dtype=np.float64)
        ids = rng.choice(np.arange(1, num_items), size=n, p=w / w.sum()).astype(np.uint64)
    else:
        ids = rng.integers(1, num_items, size=n).astype(np.uint64)
    return ptr, ids


def make_pair(pkg, O, kind, num_items, T, D, loss="hinge", optimizer="adagrad", variant="normal", lr=0.05, l2=1e-4,
              epochs=1, threads=1, parallelism="asynchronous", seed=bytes(range(1, 17)), scale=None, exact=False):
    """Builds the GPU model and an oracle model holding bit-identical parameters, optimizer state and rng."""
    H = pkg.lstm.Hyperparameters if kind == "lstm" else pkg.ewma.Hyperparameters
    h = (H(num_items, T).embedding_dim(D).learning_rate(lr).l2_penalty(l2).loss(LOSSES[loss])
         .optimizer(OPTS[optimizer]).num_epochs(epochs).num_threads(threads).parallelism(PARS[parallelism])
         .from_seed(seed))
    if kind == "lstm":
        h = h.lstm_variant(VARIANTS[variant])
    if exact:
        h = h.exact_arithmetic()
    gm = h.build()
    om = O.OracleModel(kind, num_items, T, embedding_dim=D, learning_rate=lr, l2_penalty=l2, lstm_variant=variant,
                       loss=loss, optimizer=optimizer, parallelism=parallelism, num_threads=threads,
                       num_epochs=epochs, seed=seed)
    if scale is not None:  # larger embeddings make every term of the gradient matter
        r = np.random.default_rng(123)
        gm.set_parameter("item_embeddings", (r.standard_normal(num_items * D) * scale).astype(np.float32))
        gm.set_parameter("item_biases", (r.standard_normal(num_items) * scale).astype(np.float32))
        if kind == "ewma":
            gm.set_parameter("alpha", (r.standard_normal(D) * 0.5).astype(np.float32))
    sync_params(gm, om)
    om.rng_state = gm.rng_state
    om.num_updates = gm.num_updates
    return gm, om


def sync_params(gm, om):
    for name in om.param_names():
        om.param(name)[:] = gm.get_parameter(name)


def state_names(om, optimizer):
    names = []
    for n in om.param_names():
        names.append(n)
        names.append(n + ".s1")
        if optimizer == "adam":
            names.append(n + ".s2")
    return names


def max_abs_diff(gm, om, names):
    out = {}
    for n in names:
        a, b = gm.get_parameter(n), om.param(n)
        out[n] = float(np.max(np.abs(a - b))) if len(a) else 0.0
    return out
