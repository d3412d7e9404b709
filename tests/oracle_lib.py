"""ctypes wrapper around the CPU oracle (oracle/build/liboracle.so).

Test infrastructure only: nothing under the product package imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

LOSS = {"bpr": 0, "hinge": 1, "warp": 2}
OPT = {"adagrad": 0, "adam": 1}
PAR = {"asynchronous": 0, "synchronous": 1}
VARIANT = {"normal": 0, "coupled": 1}
MODEL = {"lstm": 0, "ewma": 1}

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)


class Hyper(C.Structure):
    _fields_ = [
        ("model", C.c_int),
        ("num_items", C.c_size_t),
        ("max_sequence_length", C.c_size_t),
        ("embedding_dim", C.c_size_t),
        ("learning_rate", C.c_float),
        ("l2_penalty", C.c_float),
        ("lstm_variant", C.c_int),
        ("loss", C.c_int),
        ("optimizer", C.c_int),
        ("parallelism", C.c_int),
        ("num_threads", C.c_int),
        ("num_epochs", C.c_int),
        ("seed", C.c_uint8 * 16),
    ]


class Rng(C.Structure):
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32), ("z", C.c_uint32), ("w", C.c_uint32)]


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(ROOT, "oracle", "build", "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.sbo_model_new.restype = C.c_void_p
    L.sbo_model_new.argtypes = [C.POINTER(Hyper)]
    L.sbo_model_free.argtypes = [C.c_void_p]
    L.sbo_model_param.restype = f32p
    L.sbo_model_param.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_size_t)]
    L.sbo_model_num_updates.restype = u64p
    L.sbo_model_num_updates.argtypes = [C.c_void_p]
    L.sbo_model_rng.restype = C.POINTER(Rng)
    L.sbo_model_rng.argtypes = [C.c_void_p]
    L.sbo_fit.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t, f32p]
    L.sbo_step.restype = C.c_float
    L.sbo_step.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_uint64, C.c_uint64, u32p, C.c_int, f32p, u32p]
    L.sbo_loss_only.restype = C.c_double
    L.sbo_loss_only.argtypes = [C.c_void_p, u64p, C.c_size_t, u32p]
    L.sbo_last_sparse_grads.restype = C.c_size_t
    L.sbo_last_sparse_grads.argtypes = [C.c_void_p, C.POINTER(u32p), C.POINTER(f32p), C.POINTER(u32p),
                                        C.POINTER(f32p), C.POINTER(C.c_size_t)]
    L.sbo_user_representation.argtypes = [C.c_void_p, u64p, C.c_size_t, f32p]
    L.sbo_predict.argtypes = [C.c_void_p, f32p, u64p, C.c_size_t, f32p]
    L.sbo_mrr_score.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t, f32p]
    L.sbo_compress.argtypes = [u64p, u64p, u64p, C.c_size_t, C.c_size_t, u64p, u64p, u64p]
    L.sbo_chunks.restype = C.c_size_t
    L.sbo_chunks.argtypes = [C.c_size_t, C.c_size_t, u64p, u64p]
    L.sbo_subsequences.restype = C.c_size_t
    L.sbo_subsequences.argtypes = [u64p, C.c_size_t, C.c_size_t, u64p, u32p]
    L.sbo_user_based_split.argtypes = [u64p, C.c_size_t, C.POINTER(Rng), C.c_float, u8p]
    L.sbo_siphash24_u64.restype = C.c_uint64
    L.sbo_siphash24_u64.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    L.sbo_rng_from_seed.argtypes = [C.POINTER(Rng), u8p]
    L.sbo_rng_next_u32.restype = C.c_uint32
    L.sbo_rng_next_u32.argtypes = [C.POINTER(Rng)]
    L.sbo_rng_gen_range.restype = C.c_uint64
    L.sbo_rng_gen_range.argtypes = [C.POINTER(Rng), C.c_uint64, C.c_uint64]
    L.sbo_shuffle_u32.argtypes = [C.POINTER(Rng), u32p, C.c_size_t]
    L.sbo_draw_item.restype = C.c_uint32
    L.sbo_draw_item.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
    _LIB = L
    return L


def _p(a, t):
    return a.ctypes.data_as(t)


def make_rng(seed_bytes):
    r = Rng()
    s = (C.c_uint8 * 16)(*seed_bytes)
    lib().sbo_rng_from_seed(C.byref(r), s)
    return r


def compress(users, items, ts, num_users):
    users = np.ascontiguousarray(users, dtype=np.uint64)
    items = np.ascontiguousarray(items, dtype=np.uint64)
    ts = np.ascontiguousarray(ts, dtype=np.uint64)
    nnz = len(users)
    up = np.zeros(num_users + 1, dtype=np.uint64)
    ii = np.zeros(max(nnz, 1), dtype=np.uint64)
    tt = np.zeros(max(nnz, 1), dtype=np.uint64)
    rc = lib().sbo_compress(_p(users, u64p), _p(items, u64p), _p(ts, u64p), nnz, num_users, _p(up, u64p), _p(ii, u64p),
                            _p(tt, u64p))
    if rc:
        raise ValueError("sbo_compress rc=%d" % rc)
    return up, ii[:nnz], tt[:nnz]


def chunks(length, chunk_size):
    n = lib().sbo_chunks(length, chunk_size, None, None)
    st = np.zeros(max(n, 1), dtype=np.uint64)
    ln = np.zeros(max(n, 1), dtype=np.uint64)
    lib().sbo_chunks(length, chunk_size, _p(st, u64p), _p(ln, u64p))
    return [(int(st[i]), int(ln[i])) for i in range(n)]


def subsequences(user_ptr, max_len):
    user_ptr = np.ascontiguousarray(user_ptr, dtype=np.uint64)
    nu = len(user_ptr) - 1
    n = lib().sbo_subsequences(_p(user_ptr, u64p), nu, max_len, None, None)
    st = np.zeros(max(n, 1), dtype=np.uint64)
    ln = np.zeros(max(n, 1), dtype=np.uint32)
    lib().sbo_subsequences(_p(user_ptr, u64p), nu, max_len, _p(st, u64p), _p(ln, u32p))
    return st[:n], ln[:n]


def user_based_split(users, seed_bytes, test_fraction, rng=None):
    users = np.ascontiguousarray(users, dtype=np.uint64)
    r = rng if rng is not None else make_rng(seed_bytes)
    out = np.zeros(len(users), dtype=np.uint8)
    lib().sbo_user_based_split(_p(users, u64p), len(users), C.byref(r), test_fraction, _p(out, u8p))
    return out.astype(bool), r


class OracleModel:
    def __init__(self, model, num_items, max_sequence_length, embedding_dim=16, learning_rate=0.01, l2_penalty=0.0,
                 lstm_variant="coupled", loss="bpr", optimizer="adam", parallelism="synchronous", num_threads=1,
                 num_epochs=10, seed=bytes([42] * 16)):
        h = Hyper()
        h.model = MODEL[model]
        h.num_items = num_items
        h.max_sequence_length = max_sequence_length
        h.embedding_dim = embedding_dim
        h.learning_rate = learning_rate
        h.l2_penalty = l2_penalty
        h.lstm_variant = VARIANT[lstm_variant]
        h.loss = LOSS[loss]
        h.optimizer = OPT[optimizer]
        h.parallelism = PAR[parallelism]
        h.num_threads = num_threads
        h.num_epochs = num_epochs
        for i in range(16):
            h.seed[i] = seed[i]
        self.h = h
        self.kind = model
        self.N, self.T, self.D = num_items, max_sequence_length, embedding_dim
        self.ptr = lib().sbo_model_new(C.byref(h))

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().sbo_model_free(self.ptr)
            self.ptr = None

    def param(self, name):
        """numpy VIEW of a parameter blob (writes go straight into the oracle model)."""
        n = C.c_size_t()
        p = lib().sbo_model_param(self.ptr, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def param_names(self):
        names = ["item_embeddings", "item_biases"]
        names += ["lstm_weights", "lstm_biases"] if self.kind == "lstm" else ["alpha"]
        return names

    @property
    def num_updates(self):
        return lib().sbo_model_num_updates(self.ptr)[0]

    @num_updates.setter
    def num_updates(self, v):
        lib().sbo_model_num_updates(self.ptr)[0] = v

    @property
    def rng_state(self):
        r = lib().sbo_model_rng(self.ptr)[0]
        return (r.x, r.y, r.z, r.w)

    @rng_state.setter
    def rng_state(self, v):
        r = lib().sbo_model_rng(self.ptr)
        r[0].x, r[0].y, r[0].z, r[0].w = v

    def fit(self, user_ptr, item_ids):
        user_ptr = np.ascontiguousarray(user_ptr, dtype=np.uint64)
        item_ids = np.ascontiguousarray(item_ids, dtype=np.uint64)
        loss = C.c_float()
        rc = lib().sbo_fit(self.ptr, _p(user_ptr, u64p), _p(item_ids, u64p), len(user_ptr) - 1, C.byref(loss))
        return rc, loss.value

    def step(self, ids, key=0, step=0, apply=True, forced_negatives=None):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        negs = np.zeros(len(ids) - 1, dtype=np.uint32)
        n = C.c_size_t()
        lib().sbo_model_param(self.ptr, (b"lstm_weights" if self.kind == "lstm" else b"alpha"), C.byref(n))
        nd = n.value + (4 * self.D if self.kind == "lstm" else 0)
        dg = np.zeros(nd, dtype=np.float32)
        fn = None
        if forced_negatives is not None:
            fn_arr = np.ascontiguousarray(forced_negatives, dtype=np.uint32)
            fn = _p(fn_arr, u32p)
        l = lib().sbo_step(self.ptr, _p(ids, u64p), len(ids), key, step, _p(negs, u32p), 1 if apply else 0,
                           _p(dg, f32p), fn)
        return l, negs, dg

    def loss_only(self, ids, negatives):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        negatives = np.ascontiguousarray(negatives, dtype=np.uint32)
        return lib().sbo_loss_only(self.ptr, _p(ids, u64p), len(ids), _p(negatives, u32p))

    def last_sparse_grads(self):
        rows, grads, brows, bgrads = u32p(), f32p(), u32p(), f32p()
        nb = C.c_size_t()
        n = lib().sbo_last_sparse_grads(self.ptr, C.byref(rows), C.byref(grads), C.byref(brows), C.byref(bgrads),
                                        C.byref(nb))
        r = np.ctypeslib.as_array(rows, shape=(n,)).copy()
        g = np.ctypeslib.as_array(grads, shape=(n, self.D)).copy()
        br = np.ctypeslib.as_array(brows, shape=(nb.value,)).copy()
        bg = np.ctypeslib.as_array(bgrads, shape=(nb.value,)).copy()
        return r, g, br, bg

    def user_representation(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = np.zeros(self.D, dtype=np.float32)
        rc = lib().sbo_user_representation(self.ptr, _p(ids, u64p), len(ids), _p(out, f32p))
        return rc, out

    def predict(self, user, ids):
        user = np.ascontiguousarray(user, dtype=np.float32)
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = np.zeros(len(ids), dtype=np.float32)
        rc = lib().sbo_predict(self.ptr, _p(user, f32p), _p(ids, u64p), len(ids), _p(out, f32p))
        return rc, out

    def mrr_score(self, user_ptr, item_ids):
        user_ptr = np.ascontiguousarray(user_ptr, dtype=np.uint64)
        item_ids = np.ascontiguousarray(item_ids, dtype=np.uint64)
        out = C.c_float()
        rc = lib().sbo_mrr_score(self.ptr, _p(user_ptr, u64p), _p(item_ids, u64p), len(user_ptr) - 1, C.byref(out))
        return rc, out.value
