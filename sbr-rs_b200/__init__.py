"""Host-side mirror of the sbr-rs public API for the fit()/predict() path, over the C ABI of libsbr_b200.so.

This module contains NO compute: every method forwards to an `extern "C"` entry point declared in
include/sbr_b200.h (ctypes only; no torch, no numpy math).  Names follow the reference:

    sbr::data::{Interaction, Interactions, CompressedInteractions}      src/data.rs
    sbr::models::{Loss, Optimizer, Parallelism}                         src/models/mod.rs:16-41
    sbr::models::lstm::{Hyperparameters, LSTMVariant, ImplicitLSTMModel} src/models/lstm.rs
    sbr::models::ewma::{Hyperparameters, ImplicitEWMAModel}             src/models/ewma.rs
    sbr::evaluation::mrr_score                                          src/evaluation.rs:12-48
    sbr::{FittingError, PredictionError}                                src/lib.rs:84-97

The directory name `sbr-rs_b200` is not an importable identifier; load it with `load_package()` from
`__graft_entry__.py` (registers it as module `sbr_rs_b200`).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsbr_b200.so")

u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)

SBR_OK, SBR_ERR_NO_INTERACTIONS, SBR_ERR_INVALID_PREDICTION, SBR_ERR_INVALID_ARGUMENT = 0, 1, 2, 3
SBR_ERR_CUDA, SBR_ERR_NCCL, SBR_ERR_UNSUPPORTED = 4, 5, 6

# every symbol include/sbr_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "sbr_last_error_string", "sbr_device_count", "sbr_set_device",
    "sbr_compressed_from_triplets", "sbr_compressed_from_triplets_device", "sbr_compressed_from_csr", "sbr_compressed_borrow_csr", "sbr_compressed_num_users", "sbr_compressed_num_items",
    "sbr_compressed_len", "sbr_compressed_borrow", "sbr_compressed_user_chunks", "sbr_compressed_upload", "sbr_host_schedule", "sbr_host_master_schedule", "sbr_fit_plan_read_schedule", "sbr_user_based_split", "sbr_train_test_split",
    "sbr_compressed_free",
    "sbr_lstm_hyperparameters_new", "sbr_ewma_hyperparameters_new", "sbr_hyper_learning_rate", "sbr_hyper_l2_penalty",
    "sbr_hyper_embedding_dim", "sbr_hyper_num_epochs", "sbr_hyper_loss", "sbr_hyper_lstm_variant",
    "sbr_hyper_num_threads", "sbr_hyper_parallelism", "sbr_hyper_from_seed", "sbr_hyper_optimizer", "sbr_hyper_free",
    "sbr_hyper_build", "sbr_lstm_hyperparameters_random", "sbr_ewma_hyperparameters_random", "sbr_hyper_exact_arithmetic",
    "sbr_hyper_get_values", "sbr_model_get_hyper_values", "sbr_model_save", "sbr_model_load", "sbr_model_restore",
    "sbr_model_fit", "sbr_model_set_num_threads", "sbr_model_user_representation", "sbr_model_user_representations", "sbr_model_predict",
    "sbr_model_mrr_score", "sbr_model_gather_rows", "sbr_model_gather_rows_timed", "sbr_model_embedding_dim", "sbr_model_num_items",
    "sbr_model_parameter_len", "sbr_model_get_parameter", "sbr_model_set_parameter", "sbr_model_get_num_updates",
    "sbr_model_set_num_updates", "sbr_model_get_rng_state", "sbr_model_set_rng_state", "sbr_model_free",
    "sbr_hyper_shard", "sbr_hyper_virtual_shards", "sbr_model_ipc_handle_size", "sbr_model_ipc_export",
    "sbr_model_ipc_attach", "sbr_model_replica_sync", "sbr_dist_unique_id", "sbr_dist_init", "sbr_dist_finalize",
    "sbr_fit_plan_create", "sbr_fit_plan_run", "sbr_fit_plan_stats", "sbr_fit_plan_free", "sbr_model_last_fit_stats",
]


class HyperValues(C.Structure):
    """sbr_hyper_values: the public fields of Hyperparameters (lstm.rs:39-52) as plain data"""
    _fields_ = [
        ("model", C.c_int32), ("lstm_variant", C.c_int32), ("loss", C.c_int32), ("optimizer", C.c_int32),
        ("parallelism", C.c_int32), ("exact_arithmetic", C.c_int32),
        ("num_items", C.c_uint64), ("max_sequence_length", C.c_uint64), ("embedding_dim", C.c_uint64),
        ("num_threads", C.c_uint64), ("num_epochs", C.c_uint64),
        ("learning_rate", C.c_float), ("l2_penalty", C.c_float), ("seed", C.c_uint8 * 16),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "seed"}
        d["seed"] = bytes(self.seed)
        return d


class FitStats(C.Structure):
    _fields_ = [
        ("steps", C.c_uint64), ("timesteps", C.c_uint64), ("partitions", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("train_kernel_ms", C.c_double), ("total_device_ms", C.c_double), ("host_prepare_ms", C.c_double),
        ("upload_ms", C.c_double), ("kernel", C.c_char * 64),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["kernel"] = d["kernel"].decode()
        return d


class FittingError(Exception):
    """lib.rs:93-97"""


class NoInteractions(FittingError):
    pass


class PredictionError(Exception):
    """lib.rs:85-89"""


class InvalidPredictionValue(PredictionError):
    pass


class SbrError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("sbr status %d: %s" % (status, msg))
        self.status = status


_lib = None


def lib():
    """Load libsbr_b200.so.  Fails loudly if the CUDA extension has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libsbr_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C sbr-rs_b200`); this package has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.sbr_last_error_string.restype = C.c_char_p
    L.sbr_compressed_from_triplets.argtypes = [u64p, u64p, u64p, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(vp)]
    L.sbr_compressed_from_triplets_device.argtypes = [u64p, u64p, u64p, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(vp)]
    L.sbr_compressed_from_csr.argtypes = [u64p, u64p, u64p, C.c_size_t, C.c_size_t, C.POINTER(vp)]
    L.sbr_compressed_borrow_csr.argtypes = [u64p, u64p, u64p, C.c_size_t, C.c_size_t, C.POINTER(vp)]
    for f in ("num_users", "num_items", "len"):
        getattr(L, "sbr_compressed_" + f).restype = C.c_size_t
        getattr(L, "sbr_compressed_" + f).argtypes = [vp]
    L.sbr_compressed_borrow.argtypes = [vp, C.POINTER(u64p), C.POINTER(u64p), C.POINTER(u64p)]
    L.sbr_compressed_user_chunks.argtypes = [vp, C.c_size_t, C.c_size_t, u64p, u64p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.sbr_compressed_upload.argtypes = [vp]
    L.sbr_user_based_split.argtypes = [u64p, C.c_size_t, C.POINTER(C.c_uint32), C.c_float, C.POINTER(C.c_uint8)]
    L.sbr_train_test_split.argtypes = [C.c_size_t, C.POINTER(C.c_uint32), C.c_float, u64p, C.POINTER(C.c_size_t)]
    L.sbr_host_schedule.argtypes = [vp, C.c_size_t, C.POINTER(C.c_uint32), u64p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_size_t,
                                    C.POINTER(C.c_size_t)]
    L.sbr_host_master_schedule.argtypes = [C.POINTER(C.c_uint32), C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint32), u64p]
    L.sbr_fit_plan_read_schedule.argtypes = [vp, u64p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_size_t, C.POINTER(C.c_size_t),
                                             C.POINTER(C.c_size_t)]
    L.sbr_compressed_free.argtypes = [vp]
    L.sbr_lstm_hyperparameters_new.restype = vp
    L.sbr_lstm_hyperparameters_new.argtypes = [C.c_size_t, C.c_size_t]
    L.sbr_ewma_hyperparameters_new.restype = vp
    L.sbr_ewma_hyperparameters_new.argtypes = [C.c_size_t, C.c_size_t]
    L.sbr_hyper_learning_rate.argtypes = [vp, C.c_float]
    L.sbr_hyper_l2_penalty.argtypes = [vp, C.c_float]
    L.sbr_hyper_embedding_dim.argtypes = [vp, C.c_size_t]
    L.sbr_hyper_num_epochs.argtypes = [vp, C.c_size_t]
    L.sbr_hyper_loss.argtypes = [vp, C.c_int]
    L.sbr_hyper_lstm_variant.argtypes = [vp, C.c_int]
    L.sbr_hyper_num_threads.argtypes = [vp, C.c_size_t]
    L.sbr_hyper_parallelism.argtypes = [vp, C.c_int]
    L.sbr_hyper_from_seed.argtypes = [vp, u8p]
    L.sbr_hyper_optimizer.argtypes = [vp, C.c_int]
    L.sbr_hyper_free.argtypes = [vp]
    L.sbr_hyper_build.argtypes = [vp, C.POINTER(vp)]
    for f in ("sbr_lstm_hyperparameters_random", "sbr_ewma_hyperparameters_random"):
        getattr(L, f).restype = vp
        getattr(L, f).argtypes = [C.c_size_t, C.POINTER(C.c_uint32)]
    L.sbr_hyper_exact_arithmetic.argtypes = [vp, C.c_int]
    L.sbr_hyper_get_values.argtypes = [vp, C.POINTER(HyperValues)]
    L.sbr_model_get_hyper_values.argtypes = [vp, C.POINTER(HyperValues)]
    L.sbr_model_save.argtypes = [vp, C.c_char_p]
    L.sbr_model_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.sbr_model_restore.argtypes = [vp, C.c_char_p]
    L.sbr_model_fit.argtypes = [vp, vp, f32p]
    L.sbr_model_user_representation.argtypes = [vp, u64p, C.c_size_t, f32p]
    L.sbr_model_user_representations.argtypes = [vp, u64p, u64p, C.c_size_t, f32p]
    L.sbr_model_predict.argtypes = [vp, f32p, u64p, C.c_size_t, f32p]
    L.sbr_model_mrr_score.argtypes = [vp, vp, f32p]
    L.sbr_model_gather_rows.argtypes = [vp, u64p, C.c_size_t, f32p]
    L.sbr_model_gather_rows_timed.argtypes = [vp, u64p, C.c_size_t, f32p, C.c_int, C.POINTER(C.c_double)]
    L.sbr_model_embedding_dim.restype = C.c_size_t
    L.sbr_model_embedding_dim.argtypes = [vp]
    L.sbr_model_num_items.restype = C.c_size_t
    L.sbr_model_num_items.argtypes = [vp]
    L.sbr_model_parameter_len.argtypes = [vp, C.c_char_p, C.POINTER(C.c_size_t)]
    L.sbr_model_get_parameter.argtypes = [vp, C.c_char_p, f32p, C.c_size_t]
    L.sbr_model_set_parameter.argtypes = [vp, C.c_char_p, f32p, C.c_size_t]
    L.sbr_model_get_num_updates.argtypes = [vp, u64p]
    L.sbr_model_set_num_updates.argtypes = [vp, C.c_uint64]
    L.sbr_model_get_rng_state.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.sbr_model_set_rng_state.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.sbr_model_free.argtypes = [vp]
    L.sbr_hyper_shard.argtypes = [vp, C.c_int, C.c_int]
    L.sbr_hyper_virtual_shards.argtypes = [vp, C.c_int]
    L.sbr_model_ipc_handle_size.restype = C.c_size_t
    L.sbr_model_ipc_export.argtypes = [vp, C.c_void_p]
    L.sbr_model_ipc_attach.argtypes = [vp, C.c_void_p]
    L.sbr_model_replica_sync.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.sbr_model_set_num_threads.argtypes = [vp, C.c_size_t]
    L.sbr_dist_unique_id.argtypes = [u8p]
    L.sbr_dist_init.argtypes = [C.c_int, C.c_int, u8p]
    L.sbr_dist_finalize.restype = None
    L.sbr_fit_plan_create.argtypes = [vp, vp, C.POINTER(vp)]
    L.sbr_fit_plan_run.argtypes = [vp, f32p]
    L.sbr_fit_plan_stats.argtypes = [vp, C.POINTER(FitStats)]
    L.sbr_fit_plan_free.argtypes = [vp]
    L.sbr_model_last_fit_stats.argtypes = [vp, C.POINTER(FitStats)]
    _lib = L
    return L


def _check(status):
    if status == SBR_OK:
        return
    msg = lib().sbr_last_error_string().decode()
    if status == SBR_ERR_NO_INTERACTIONS:
        raise NoInteractions(msg)
    if status == SBR_ERR_INVALID_PREDICTION:
        raise InvalidPredictionValue(msg)
    raise SbrError(status, msg)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p(a, t):
    return a.ctypes.data_as(t)


def device_count():
    return lib().sbr_device_count()


def set_device(i):
    _check(lib().sbr_set_device(i))


def dist_unique_id():
    buf = (C.c_uint8 * 128)()
    _check(lib().sbr_dist_unique_id(buf))
    return bytes(buf)


def dist_init(rank, world, unique_id):
    buf = (C.c_uint8 * 128)(*unique_id)
    _check(lib().sbr_dist_init(rank, world, buf))


def dist_finalize():
    lib().sbr_dist_finalize()


# ----------------------------------------------------------------------------------------------- data.rs ----
class Interaction:
    """data.rs:16-51"""
    __slots__ = ("_u", "_i", "_t")

    def __init__(self, user_id, item_id, timestamp):
        self._u, self._i, self._t = int(user_id), int(item_id), int(timestamp)

    def user_id(self):
        return self._u

    def item_id(self):
        return self._i

    def timestamp(self):
        return self._t

    def weight(self):
        return 1.0


class Interactions:
    """data.rs:91-211 (AoS in the reference; kept as three columns here)."""

    def __init__(self, num_users, num_items):
        self._num_users, self._num_items = int(num_users), int(num_items)
        self._u, self._i, self._t = [], [], []

    @classmethod
    def from_interactions(cls, interactions):
        """impl From<Vec<Interaction>> (data.rs:200-211): num_users = max+1, num_items = max+1."""
        interactions = list(interactions)
        self = cls(max(x.user_id() for x in interactions) + 1, max(x.item_id() for x in interactions) + 1)
        for x in interactions:
            self.push(x)
        return self

    @classmethod
    def from_arrays(cls, user_ids, item_ids, timestamps, num_users=None, num_items=None):
        user_ids, item_ids, timestamps = _u64(user_ids), _u64(item_ids), _u64(timestamps)
        self = cls(int(user_ids.max()) + 1 if num_users is None else num_users,
                   int(item_ids.max()) + 1 if num_items is None else num_items)
        self._u, self._i, self._t = user_ids, item_ids, timestamps
        return self

    def push(self, x):
        self._u = list(self._u); self._i = list(self._i); self._t = list(self._t)
        self._u.append(x.user_id()); self._i.append(x.item_id()); self._t.append(x.timestamp())

    def data(self):
        return [Interaction(u, i, t) for u, i, t in zip(self._u, self._i, self._t)]

    def __len__(self):
        return len(self._u)

    def len(self):
        return len(self._u)

    def is_empty(self):
        return len(self._u) == 0

    def num_users(self):
        return self._num_users

    def num_items(self):
        return self._num_items

    def shape(self):
        return (self._num_users, self._num_items)

    def split_by(self, mask):
        """data.rs:149-172 with a precomputed predicate column."""
        mask = np.asarray(mask, dtype=bool)
        u, i, t = _u64(self._u), _u64(self._i), _u64(self._t)
        a = Interactions.from_arrays(u[mask], i[mask], t[mask], self._num_users, self._num_items)
        b = Interactions.from_arrays(u[~mask], i[~mask], t[~mask], self._num_users, self._num_items)
        return a, b

    def to_compressed(self):
        """data.rs:180-182"""
        return CompressedInteractions._from_triplets(_u64(self._u), _u64(self._i), _u64(self._t), self._num_users,
                                                     self._num_items)


class CompressedInteractions:
    """data.rs:228-329.  Owns an sbr_compressed handle."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def _from_triplets(cls, u, i, t, num_users, num_items, device=False):
        """device=True: the CSR is built on the GPU (sbr_compressed_from_triplets_device) and stays resident."""
        h = C.c_void_p()
        fn = lib().sbr_compressed_from_triplets_device if device else lib().sbr_compressed_from_triplets
        _check(fn(_p(u, u64p), _p(i, u64p), _p(t, u64p), len(u), num_users, num_items, C.byref(h)))
        return cls(h)

    @classmethod
    def from_csr(cls, user_pointers, item_ids, timestamps=None, num_items=None, borrow=False):
        """borrow=True: zero-copy view of the caller's arrays (sbr_compressed_borrow_csr); they are kept alive here."""
        up, ii = _u64(user_pointers), _u64(item_ids)
        tt = _u64(timestamps) if timestamps is not None else None
        if num_items is None:
            num_items = int(ii.max()) + 1 if len(ii) else 0
        h = C.c_void_p()
        fn = lib().sbr_compressed_borrow_csr if borrow else lib().sbr_compressed_from_csr
        _check(fn(_p(up, u64p), _p(ii, u64p), _p(tt, u64p) if tt is not None else None, len(up) - 1, num_items, C.byref(h)))
        self = cls(h)
        if borrow:
            self._keep = (up, ii, tt)
        return self

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sbr_compressed_free(self._h)
            self._h = None

    def num_users(self):
        return lib().sbr_compressed_num_users(self._h)

    def num_items(self):
        return lib().sbr_compressed_num_items(self._h)

    def shape(self):
        return (self.num_users(), self.num_items())

    def __len__(self):
        return lib().sbr_compressed_len(self._h)

    def arrays(self):
        """(user_pointers, item_ids, timestamps) copies of the CSR fields (data.rs:231-233)."""
        up, ii, tt = u64p(), u64p(), u64p()
        _check(lib().sbr_compressed_borrow(self._h, C.byref(up), C.byref(ii), C.byref(tt)))
        nu, nnz = self.num_users(), len(self)
        # a NULL pointer (a CSR borrowed without timestamps) reads as zeros, like sbr_compressed_from_csr(timestamps=NULL)
        f = lambda p, n: np.ctypeslib.as_array(p, shape=(n,)).copy() if (n and p) else np.zeros(n, dtype=np.uint64)
        return f(up, nu + 1), f(ii, nnz), f(tt, nnz)

    def to_interactions(self):
        """data.rs:308-328"""
        up, ii, tt = self.arrays()
        users = np.repeat(np.arange(self.num_users(), dtype=np.uint64), np.diff(up).astype(np.int64))
        return Interactions.from_arrays(users, ii, tt, self.num_users(), self.num_items())

    def user_chunks(self, user_id, chunk_size):
        """CompressedInteractionsUser::chunks (data.rs:363-371): list of (start, len) within the user's slice."""
        n = C.c_size_t()
        _check(lib().sbr_compressed_user_chunks(self._h, user_id, chunk_size, None, None, 0, C.byref(n)))
        st = np.zeros(max(n.value, 1), dtype=np.uint64)
        ln = np.zeros(max(n.value, 1), dtype=np.uint64)
        _check(lib().sbr_compressed_user_chunks(self._h, user_id, chunk_size, _p(st, u64p), _p(ln, u64p), n.value,
                                                C.byref(n)))
        return [(int(st[k]), int(ln[k])) for k in range(n.value)]

    def upload(self):
        _check(lib().sbr_compressed_upload(self._h))
        return self

    def host_schedule(self, max_sequence_length, rng_state):
        """sequence_model.rs:76-84 as fit() does it on the host (sbr_host_schedule; needs no device):
        returns (starts, lens, shuffled order, rng state after the shuffle)."""
        st = (C.c_uint32 * 4)(*[int(x) for x in rng_state])
        n = C.c_size_t()
        _check(lib().sbr_host_schedule(self._h, max_sequence_length, st, None, None, None, 0, C.byref(n)))
        starts = np.zeros(n.value, dtype=np.uint64); lens = np.zeros(n.value, dtype=np.uint32); order = np.zeros(n.value, dtype=np.uint32)
        _check(lib().sbr_host_schedule(self._h, max_sequence_length, st, _p(starts, u64p), lens.ctypes.data_as(C.POINTER(C.c_uint32)),
                                       order.ctypes.data_as(C.POINTER(C.c_uint32)), n.value, C.byref(n)))
        return starts, lens, order, tuple(int(x) for x in st)


def host_master_schedule(rng_state, nsub, partitions, threads):
    """sequence_model.rs:84,97 as fit() does it (sbr_host_master_schedule; host only): (shuffled indices, partition keys, rng state after)."""
    st = (C.c_uint32 * 4)(*[int(x) for x in rng_state])
    order = np.zeros(nsub, dtype=np.uint32); keys = np.zeros(max(partitions, 1), dtype=np.uint64)
    _check(lib().sbr_host_master_schedule(st, nsub, partitions, threads, order.ctypes.data_as(C.POINTER(C.c_uint32)), _p(keys, u64p)))
    return order, keys[:partitions], tuple(int(x) for x in st)


def user_based_split(user_ids, rng_state, test_fraction):
    """data.rs:69-88: boolean mask is_train per interaction (no user in both sets) and the rng state afterwards."""
    u = _u64(user_ids)
    st = (C.c_uint32 * 4)(*[int(x) for x in rng_state])
    out = np.zeros(len(u), dtype=np.uint8)
    _check(lib().sbr_user_based_split(_p(u, u64p), len(u), st, float(test_fraction), out.ctypes.data_as(C.POINTER(C.c_uint8))))
    return out.astype(bool), tuple(int(x) for x in st)


def train_test_split(nnz, rng_state, test_fraction):
    """data.rs:54-64: (train_indices, test_indices) into the original interactions and the rng state afterwards."""
    st = (C.c_uint32 * 4)(*[int(x) for x in rng_state])
    perm = np.zeros(int(nnz), dtype=np.uint64)
    nt = C.c_size_t()
    _check(lib().sbr_train_test_split(int(nnz), st, float(test_fraction), _p(perm, u64p), C.byref(nt)))
    return perm[nt.value:], perm[:nt.value], tuple(int(x) for x in st)


# ------------------------------------------------------------------------------------------------ models ----
class Loss:
    BPR, Hinge, WARP = 0, 1, 2


class Optimizer:
    Adagrad, Adam = 0, 1


class Parallelism:
    Asynchronous, Synchronous = 0, 1


class LSTMVariant:
    Normal, Coupled = 0, 1


class ImplicitUser:
    """models/mod.rs:9-12"""

    def __init__(self, user_embedding):
        self.user_embedding = np.ascontiguousarray(user_embedding, dtype=np.float32)


class _Hyperparameters:
    _KIND = None

    def __init__(self, num_items, max_sequence_length):
        new = lib().sbr_lstm_hyperparameters_new if self._KIND == "lstm" else lib().sbr_ewma_hyperparameters_new
        self._h = C.c_void_p(new(num_items, max_sequence_length))
        if not self._h:
            raise MemoryError()

    @classmethod
    def random(cls, num_items, rng_state):
        """Hyperparameters::random(num_items, rng) (lstm.rs:141-172): returns (hyperparameters, advanced rng state)."""
        fn = lib().sbr_lstm_hyperparameters_random if cls._KIND == "lstm" else lib().sbr_ewma_hyperparameters_random
        st = (C.c_uint32 * 4)(*rng_state)
        h = fn(num_items, st)
        if not h:
            raise SbrError(SBR_ERR_INVALID_ARGUMENT, lib().sbr_last_error_string().decode())
        self = cls.__new__(cls)
        self._h = C.c_void_p(h)
        return self, tuple(st)

    def values(self):
        v = HyperValues()
        _check(lib().sbr_hyper_get_values(self._h, C.byref(v)))
        return v.as_dict()

    def exact_arithmetic(self, on=True):
        """engine knob: keep exact fp32 arithmetic (never the tf32 / bf16 tensor-core tile kernel)"""
        return self._set("sbr_hyper_exact_arithmetic", int(bool(on)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sbr_hyper_free(self._h)
            self._h = None

    def _set(self, fn, v):
        _check(getattr(lib(), fn)(self._h, v))
        return self

    def learning_rate(self, v):
        return self._set("sbr_hyper_learning_rate", float(v))

    def l2_penalty(self, v):
        return self._set("sbr_hyper_l2_penalty", float(v))

    def embedding_dim(self, v):
        return self._set("sbr_hyper_embedding_dim", int(v))

    def num_epochs(self, v):
        return self._set("sbr_hyper_num_epochs", int(v))

    def loss(self, v):
        return self._set("sbr_hyper_loss", int(v))

    def num_threads(self, v):
        return self._set("sbr_hyper_num_threads", int(v))

    def parallelism(self, v):
        return self._set("sbr_hyper_parallelism", int(v))

    def optimizer(self, v):
        return self._set("sbr_hyper_optimizer", int(v))

    def shard(self, rank, world):
        """one process per GPU: this rank owns item rows with id % world == rank (sbr_hyper_shard)"""
        _check(lib().sbr_hyper_shard(self._h, int(rank), int(world)))
        return self

    def virtual_shards(self, shards):
        _check(lib().sbr_hyper_virtual_shards(self._h, int(shards)))
        return self

    def from_seed(self, seed):
        s = (C.c_uint8 * 16)(*bytes(seed))
        _check(lib().sbr_hyper_from_seed(self._h, s))
        return self

    def build(self):
        m = C.c_void_p()
        h, self._h = self._h, None  # build(self) consumes
        _check(lib().sbr_hyper_build(h, C.byref(m)))
        return self._MODEL(m)


class _Model:
    def __init__(self, handle):
        self._m = handle

    def __del__(self):
        if getattr(self, "_m", None) and _lib is not None:
            _lib.sbr_model_free(self._m)
            self._m = None

    # -- reference API --
    def fit(self, interactions):
        loss = C.c_float()
        _check(lib().sbr_model_fit(self._m, interactions._h, C.byref(loss)))
        return loss.value

    def user_representation(self, item_ids):
        ids = _u64(item_ids)
        out = np.zeros(self.embedding_dim, dtype=np.float32)
        _check(lib().sbr_model_user_representation(self._m, _p(ids, u64p), len(ids), _p(out, f32p)))
        return ImplicitUser(out)

    def predict(self, user, item_ids):
        ids = _u64(item_ids)
        out = np.zeros(len(ids), dtype=np.float32)
        _check(lib().sbr_model_predict(self._m, _p(user.user_embedding, f32p), _p(ids, u64p), len(ids), _p(out, f32p)))
        return out

    # -- extras --
    @property
    def embedding_dim(self):
        return lib().sbr_model_embedding_dim(self._m)

    @property
    def num_items(self):
        return lib().sbr_model_num_items(self._m)

    def user_representations(self, ptr, item_ids):
        ptr, ids = _u64(ptr), _u64(item_ids)
        out = np.zeros((len(ptr) - 1, self.embedding_dim), dtype=np.float32)
        _check(lib().sbr_model_user_representations(self._m, _p(ptr, u64p), _p(ids, u64p), len(ptr) - 1, _p(out, f32p)))
        return out

    def gather_rows(self, item_ids):
        ids = _u64(item_ids)
        out = np.zeros((len(ids), self.embedding_dim), dtype=np.float32)
        _check(lib().sbr_model_gather_rows(self._m, _p(ids, u64p), len(ids), _p(out, f32p)))
        return out

    def gather_rows_timed(self, item_ids, iters=5):
        """mean device time (ms) of the stand-alone gather kernel over `iters` launches (sbr_model_gather_rows_timed)"""
        ids = _u64(item_ids)
        ms = C.c_double()
        _check(lib().sbr_model_gather_rows_timed(self._m, _p(ids, u64p), len(ids), None, int(iters), C.byref(ms)))
        return ms.value

    def get_parameter(self, name):
        n = C.c_size_t()
        _check(lib().sbr_model_parameter_len(self._m, name.encode(), C.byref(n)))
        out = np.zeros(n.value, dtype=np.float32)
        _check(lib().sbr_model_get_parameter(self._m, name.encode(), _p(out, f32p), n.value))
        return out

    def set_parameter(self, name, data):
        data = np.ascontiguousarray(data, dtype=np.float32).ravel()
        _check(lib().sbr_model_set_parameter(self._m, name.encode(), _p(data, f32p), len(data)))

    @property
    def num_updates(self):
        v = C.c_uint64()
        _check(lib().sbr_model_get_num_updates(self._m, C.byref(v)))
        return v.value

    @num_updates.setter
    def num_updates(self, v):
        _check(lib().sbr_model_set_num_updates(self._m, v))

    @property
    def rng_state(self):
        st = (C.c_uint32 * 4)()
        _check(lib().sbr_model_get_rng_state(self._m, st))
        return tuple(st)

    @rng_state.setter
    def rng_state(self, v):
        st = (C.c_uint32 * 4)(*v)
        _check(lib().sbr_model_set_rng_state(self._m, st))

    # -- checkpoint: written and read by the C library (sbr_model_save / load / restore; layout in include/sbr_b200.h) --
    def hyper_values(self):
        v = HyperValues()
        _check(lib().sbr_model_get_hyper_values(self._m, C.byref(v)))
        return v.as_dict()

    def save(self, path):
        _check(lib().sbr_model_save(self._m, os.fsencode(path)))

    def restore(self, path):
        """overwrite parameters, optimizer state, rng and update counter from a checkpoint of the same shape"""
        _check(lib().sbr_model_restore(self._m, os.fsencode(path)))

    def set_num_threads(self, n):
        """sbr_model_set_num_threads: partition count of the following fits (0 = automatic)."""
        _check(lib().sbr_model_set_num_threads(self._m, int(n)))
        return self

    def replica_sync(self):
        """sbr_model_replica_sync: sum of all ranks' parameter / optimizer-state deltas applied to every replica; returns the bytes all-reduced."""
        n = C.c_size_t()
        _check(lib().sbr_model_replica_sync(self._m, C.byref(n)))
        return n.value

    def ipc_export(self):
        buf = C.create_string_buffer(lib().sbr_model_ipc_handle_size())
        _check(lib().sbr_model_ipc_export(self._m, buf))
        return buf.raw

    def ipc_attach(self, all_handles):
        """all_handles: list of per-rank blobs from ipc_export(), in rank order"""
        blob = b"".join(all_handles)
        _check(lib().sbr_model_ipc_attach(self._m, C.create_string_buffer(blob, len(blob))))

    def last_fit_stats(self):
        s = FitStats()
        _check(lib().sbr_model_last_fit_stats(self._m, C.byref(s)))
        return s.as_dict()

    def fit_plan(self, interactions):
        return FitPlan(self, interactions)


class FitPlan:
    """Device-resident schedule of one fit() (sbr_fit_plan_*): stage once, run many times."""

    def __init__(self, model, interactions):
        self._model, self._interactions = model, interactions  # keep alive
        self._p = C.c_void_p()
        _check(lib().sbr_fit_plan_create(model._m, interactions._h, C.byref(self._p)))

    def run(self):
        loss = C.c_float()
        _check(lib().sbr_fit_plan_run(self._p, C.byref(loss)))
        return loss.value

    def read_schedule(self):
        """(starts, lens, order) as staged in HBM: the device chunker's sub-sequences and the partition-major shuffled order."""
        n, no = C.c_size_t(), C.c_size_t()
        _check(lib().sbr_fit_plan_read_schedule(self._p, None, None, None, 0, C.byref(n), C.byref(no)))
        starts = np.zeros(n.value, dtype=np.uint64); lens = np.zeros(n.value, dtype=np.uint32); order = np.zeros(max(n.value, no.value), dtype=np.uint32)
        _check(lib().sbr_fit_plan_read_schedule(self._p, _p(starts, u64p), lens.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                order.ctypes.data_as(C.POINTER(C.c_uint32)), n.value, C.byref(n), C.byref(no)))
        return starts, lens, order[:no.value]

    def stats(self):
        s = FitStats()
        _check(lib().sbr_fit_plan_stats(self._p, C.byref(s)))
        return s.as_dict()

    def __del__(self):
        if getattr(self, "_p", None) and _lib is not None:
            _lib.sbr_fit_plan_free(self._p)
            self._p = None


class ImplicitLSTMModel(_Model):
    """lstm.rs:387-416"""


class ImplicitEWMAModel(_Model):
    """ewma.rs:402-429"""


class _LstmHyperparameters(_Hyperparameters):
    """lstm.rs:39-202"""
    _KIND = "lstm"
    _MODEL = ImplicitLSTMModel

    def lstm_variant(self, v):
        return self._set("sbr_hyper_lstm_variant", int(v))


class _EwmaHyperparameters(_Hyperparameters):
    """ewma.rs:45-206"""
    _KIND = "ewma"
    _MODEL = ImplicitEWMAModel


class lstm:  # namespace mirror of sbr::models::lstm
    Hyperparameters = _LstmHyperparameters
    LSTMVariant = LSTMVariant
    ImplicitLSTMModel = ImplicitLSTMModel


class ewma:  # namespace mirror of sbr::models::ewma
    Hyperparameters = _EwmaHyperparameters
    ImplicitEWMAModel = ImplicitEWMAModel


def load_model(path):
    """sbr_model_load: a new model (hyperparameters included) from a checkpoint written by Model.save()"""
    m = C.c_void_p()
    _check(lib().sbr_model_load(os.fsencode(path), C.byref(m)))
    v = HyperValues()
    _check(lib().sbr_model_get_hyper_values(m, C.byref(v)))
    return (ImplicitLSTMModel if v.model == 0 else ImplicitEWMAModel)(m)


def mrr_score(model, test):
    """evaluation.rs:12-48"""
    out = C.c_float()
    _check(lib().sbr_model_mrr_score(model._m, test._h, C.byref(out)))
    return out.value
