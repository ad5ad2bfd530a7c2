"""ctypes binding of include/hnsw_b200.h -- the host-side mirror of the reference's access-method
interface for the HNSW hot path (hnswbuild / hnswinsert / hnswbeginscan / hnswrescan / hnswgettuple
/ hnswendscan) and of pgvector's operator classes.  The reference mount holds no source
(/root/reference/README.md:1), so names follow upstream pgvector and the PostgreSQL index-AM API.

All arithmetic happens in libhnsw_b200.so's CUDA kernels.  No CPU fallback exists: a missing
library or a missing device raises HnswError.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhnsw_b200.so")
_CSRC = os.path.join(_HERE, "csrc")

HB_L2, HB_IP, HB_COSINE, HB_L1 = 0, 1, 2, 3
HB_F32, HB_F16 = 0, 1
HB_HEAPTIDS = 10

# operator class name -> (metric, dtype): what `USING hnsw (col <opclass>)` selects
OPCLASSES = {
    "vector_l2_ops": (HB_L2, HB_F32),
    "vector_ip_ops": (HB_IP, HB_F32),
    "vector_cosine_ops": (HB_COSINE, HB_F32),
    "halfvec_l2_ops": (HB_L2, HB_F16),
    "halfvec_ip_ops": (HB_IP, HB_F16),
    "halfvec_cosine_ops": (HB_COSINE, HB_F16),
    "vector_l1_ops": (HB_L1, HB_F32),
    "halfvec_l1_ops": (HB_L1, HB_F16),
}


class HnswError(RuntimeError):
    pass


class Counters(C.Structure):
    _fields_ = [("n_dist", C.c_int64), ("n_hop0", C.c_int64), ("n_hopu", C.c_int64), ("n_pair", C.c_int64),
                ("n_slow", C.c_int64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def lib_path():
    return _SO


def build_library(force=False, jobs=8):
    """Compile csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", _CSRC, "clean", "-s"])
    subprocess.check_call(["make", "-C", _CSRC, "-s", "-j%d" % jobs])
    return _SO


_lib = None

_SIGS = {
    "hb_last_error": (C.c_char_p, []),
    "hb_version": (C.c_char_p, []),
    "hb_device_count": (C.c_int, []),
    "hb_index_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_uint64]),
    "hb_index_free": (None, [C.c_void_p]),
    "hb_index_size": (C.c_int64, [C.c_void_p]),
    "hb_index_entry": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_build": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "hb_insert": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "hb_set_build_batch": (C.c_int, [C.c_void_p, C.c_int]),
    "hb_index_trim": (C.c_int, [C.c_void_p]),
    "hb_index_reserve": (C.c_int, [C.c_void_p, C.c_int64]),
    "hb_bulk_delete": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64]),
    "hb_vacuum_repair": (C.c_int64, [C.c_void_p, C.c_void_p]),
    "hb_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "hb_level_for": (C.c_int, [C.c_uint64, C.c_int64, C.c_int]),
    "hb_index_load": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32] + [C.c_void_p] * 7),
    "hb_index_upper_rows": (C.c_int64, [C.c_void_p]),
    "hb_index_load_pgvector_pages": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "hb_pgvector_pages_info": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 5),
    "hb_index_export": (C.c_int, [C.c_void_p] + [C.c_void_p] * 7),
    "hb_beginscan": (C.c_void_p, [C.c_void_p]),
    "hb_rescan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "hb_gettuple": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_endscan": (None, [C.c_void_p]),
    "hb_scan_set_iterative": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "hb_iter_begin": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64]),
    "hb_iter_next": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_iter_tuples": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hb_iter_end": (None, [C.c_void_p]),
    "hb_search_batch_filtered": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_search_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_search_batch_async": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_search_batch_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "hb_search_batch_elements": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_search_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_search_batch_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hb_search_layer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_get_counters": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "hb_get_per_query_counters": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "hb_last_search_ms": (C.c_float, [C.c_void_p]),
    "hb_distance_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
    "hb_distance_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "hb_normalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "hb_bruteforce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "hb_bruteforce_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_partition_of": (C.c_int, [C.c_int64, C.c_int]),
    "hb_partition_route": (None, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "hb_merge_topk_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_elements_to_tids_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_part_unique_id": (C.c_int, [C.c_void_p]),
    "hb_part_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_uint64, C.c_int, C.c_int, C.c_void_p]),
    "hb_part_free": (None, [C.c_void_p]),
    "hb_part_owned": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hb_part_index": (C.c_void_p, [C.c_void_p, C.c_int]),
    "hb_part_size": (C.c_int64, [C.c_void_p]),
    "hb_part_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "hb_part_get_counters": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "hb_part_build": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "hb_part_search_async": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "hb_part_search_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "hb_part_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}


def load_library():
    """dlopen libhnsw_b200.so; raises HnswError when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise HnswError("libhnsw_b200.so is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                        "There is no CPU fallback.")
    L = C.CDLL(_SO)
    for name, (res, args) in _SIGS.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _p(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def _err(L, what):
    return HnswError("%s: %s" % (what, L.hb_last_error().decode()))


def _np_dtype(dtype):
    return np.float32 if dtype == HB_F32 else np.float16


def partition_of(tid, n_partitions):
    """Partition routing of the fork-level spec: splitmix64(id) mod P."""
    return load_library().hb_partition_of(int(tid), n_partitions)


def partition_route(ids, n_partitions):
    ids = np.ascontiguousarray(ids, np.int64)
    out = np.empty(ids.shape[0], np.int32)
    load_library().hb_partition_route(_p(ids), ids.shape[0], n_partitions, _p(out))
    return out


def merge_topk_dev(device, dev_tids, dev_dist, n_parts, nq, k, dev_out_tids, dev_out_dist, stream=0):
    """Merge P x nq x k per-partition lists resident in HBM (raw device pointers)."""
    L = load_library()
    rc = L.hb_merge_topk_dev(device, _p(dev_tids), _p(dev_dist), n_parts, nq, k, _p(dev_out_tids), _p(dev_out_dist),
                             C.c_void_p(stream))
    if rc < 0:
        raise _err(L, "hb_merge_topk_dev")


class HnswIndex:
    """One HNSW index (one partition) resident on one B200.

    Mirrors `CREATE INDEX ... USING hnsw (col <opclass>) WITH (m = .., ef_construction = ..)`.
    """

    def __init__(self, dim, opclass="vector_l2_ops", m=16, ef_construction=64, capacity=1 << 20, device=0, seed=0):
        if opclass not in OPCLASSES:
            raise HnswError("operator class %r does not exist for access method hnsw" % (opclass,))
        self.metric, self.dtype = OPCLASSES[opclass]
        self.opclass, self.dim, self.m, self.efc, self.device, self.seed = opclass, dim, m, ef_construction, device, seed
        self.capacity = capacity
        self._L = load_library()
        self._h = self._L.hb_index_create(device, dim, m, ef_construction, self.metric, self.dtype, capacity, seed)
        if not self._h:
            raise _err(self._L, "hb_index_create")

    @classmethod
    def _view(cls, handle, dim, opclass, m, ef_construction, device, seed=0):
        """A non-owning view of an hb_index that belongs to someone else (an hb_part's partition)."""
        self = cls.__new__(cls)
        self.metric, self.dtype = OPCLASSES[opclass]
        self.opclass, self.dim, self.m, self.efc, self.device, self.seed = opclass, dim, m, ef_construction, device, seed
        self.capacity = None
        self._L = load_library()
        self._h = handle
        self._borrowed = True
        return self

    def close(self):
        if getattr(self, "_h", None):
            if not getattr(self, "_borrowed", False):
                self._L.hb_index_free(self._h)
            self._h = None

    __del__ = close

    # ---- helpers
    def _vecs(self, a):
        a = np.ascontiguousarray(a, _np_dtype(self.dtype))
        if a.ndim == 1:
            a = a[None, :]
        if a.shape[-1] != self.dim:
            # same wording as pgvector's CheckExpectedDim
            raise HnswError("expected %d dimensions, not %d" % (self.dim, a.shape[-1]))
        return a

    def _ck(self, rc, what):
        if rc < 0:
            raise _err(self._L, what)
        return rc

    def set_option(self, name, value):
        self._ck(self._L.hb_set_option(self._h, name.encode(), int(value)), "hb_set_option")

    @property
    def n(self):
        return int(self._L.hb_index_size(self._h))

    @property
    def entry(self):
        e, lv = C.c_int32(), C.c_int()
        self._L.hb_index_entry(self._h, C.byref(e), C.byref(lv))
        return e.value, lv.value

    # ---- ambuild / aminsert
    def build(self, vecs, heap_tids=None):
        """hnswbuild: index every row of `vecs` (heap TID = row number unless given)."""
        vecs = self._vecs(vecs)
        t = None if heap_tids is None else np.ascontiguousarray(heap_tids, np.int64)
        return self._ck(self._L.hb_build(self._h, _p(vecs), vecs.shape[0], _p(t)), "hb_build")

    def bulk_delete(self, dead_tids):
        """ambulkdelete, first pass: remove heap TIDs; returns how many were removed"""
        t = np.ascontiguousarray(dead_tids, np.int64)
        return self._ck(self._L.hb_bulk_delete(self._h, _p(t), t.size), "hb_bulk_delete")

    def vacuum_repair(self):
        """ambulkdelete, passes 2 and 3 (RepairGraph + MarkDeleted) -> (elements marked deleted, elements re-linked)"""
        rep = C.c_int64()
        marked = self._ck(self._L.hb_vacuum_repair(self._h, C.byref(rep)), "hb_vacuum_repair")
        return int(marked), int(rep.value)

    def reserve(self, capacity):
        self._ck(self._L.hb_index_reserve(self._h, capacity), "hb_index_reserve")

    def trim(self):
        """free the build-only memory (pair cache, cached neighbour distances, workspaces)"""
        self._ck(self._L.hb_index_trim(self._h), "hb_index_trim")

    def insert(self, vecs, heap_tids=None):
        """hnswinsert, batched."""
        vecs = self._vecs(vecs)
        t = None if heap_tids is None else np.ascontiguousarray(heap_tids, np.int64)
        return self._ck(self._L.hb_insert(self._h, _p(vecs), vecs.shape[0], _p(t)), "hb_insert")

    # ---- graph image
    def load_graph(self, g):
        vecs = np.ascontiguousarray(g.vecs, _np_dtype(self.dtype))
        arrs = [np.ascontiguousarray(g.level, np.uint8), np.ascontiguousarray(g.nbr0, np.int32),
                np.ascontiguousarray(g.uoff, np.int32), np.ascontiguousarray(g.nbru, np.int32),
                np.ascontiguousarray(g.ntids, np.uint8), np.ascontiguousarray(g.tids, np.int64)]
        self._ck(self._L.hb_index_load(self._h, g.n, g.upper_rows, g.entry, _p(vecs), *[_p(a) for a in arrs]),
                 "hb_index_load")

    def load_pgvector_pages(self, pages):
        """pages: bytes / uint8 array holding the index relation's 8 kB blocks (metapage first)."""
        buf = np.frombuffer(pages, np.uint8) if isinstance(pages, (bytes, bytearray)) else np.ascontiguousarray(pages, np.uint8)
        if buf.size % 8192:
            raise HnswError("index pages must be a multiple of 8192 bytes")
        self._ck(self._L.hb_index_load_pgvector_pages(self._h, _p(buf), buf.size // 8192), "hb_index_load_pgvector_pages")

    def export_graph(self):
        n, m = self.n, self.m
        ur = int(self._L.hb_index_upper_rows(self._h))

        class G:
            pass
        g = G()
        g.dim, g.m, g.efc, g.metric, g.dtype, g.n, g.upper_rows = self.dim, m, self.efc, self.metric, self.dtype, n, ur
        g.entry, g.entry_level = self.entry
        g.vecs = np.empty((n, self.dim), _np_dtype(self.dtype))
        g.level = np.empty(n, np.uint8)
        g.nbr0 = np.empty((n, 2 * m), np.int32)
        g.uoff = np.empty(n, np.int32)
        g.nbru = np.full((max(ur, 1), m), -1, np.int32)
        g.ntids = np.empty(n, np.uint8)
        g.tids = np.empty((n, HB_HEAPTIDS), np.int64)
        self._ck(self._L.hb_index_export(self._h, _p(g.vecs), _p(g.level), _p(g.nbr0), _p(g.uoff), _p(g.nbru),
                                         _p(g.ntids), _p(g.tids)), "hb_index_export")
        return g

    # ---- scans
    def beginscan(self):
        return HnswScan(self)

    def search_filtered(self, queries, k, ef_search, allowed, max_scan_tuples=20000):
        """`ORDER BY ... LIMIT k` with a filter the index cannot evaluate: `allowed` is a boolean array over heap
        TIDs (True = the tuple qualifies); scans are resumed (hnsw.iterative_scan) until k tuples pass."""
        q = self._vecs(queries)
        nq = q.shape[0]
        bits = np.packbits(np.ascontiguousarray(allowed, bool), bitorder="little")
        tids = np.empty((nq, k), np.int64)
        dist = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.int32)
        self._ck(self._L.hb_search_batch_filtered(self._h, _p(q), nq, ef_search, k, _p(bits), len(allowed), max_scan_tuples,
                                                  _p(tids), _p(dist), _p(cnt)), "hb_search_batch_filtered")
        return tids, dist, cnt

    def iterate(self, queries, ef_search=40, max_scan_tuples=20000):
        """hnsw.iterative_scan, batched: a resumable scan per query (see HnswIterator)."""
        return HnswIterator(self, queries, ef_search, max_scan_tuples)

    def search(self, queries, k=10, ef_search=40):
        """Batched `ORDER BY col <op> $1 LIMIT k`: host arrays in, host arrays out."""
        q = self._vecs(queries)
        nq = q.shape[0]
        tids = np.empty((nq, k), np.int64)
        dist = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.int32)
        self._ck(self._L.hb_search_batch(self._h, _p(q), nq, ef_search, k, _p(tids), _p(dist), _p(cnt)), "hb_search_batch")
        return tids, dist, cnt

    def search_into(self, q_ptr, nq, k, ef_search, tids_ptr, dist_ptr, cnt_ptr):
        """hb_search_batch on raw host pointers (e.g. pinned torch tensors)."""
        self._ck(self._L.hb_search_batch(self._h, C.c_void_p(q_ptr), nq, ef_search, k, C.c_void_p(tids_ptr),
                                         C.c_void_p(dist_ptr), C.c_void_p(cnt_ptr)), "hb_search_batch")

    def search_async(self, slot, q_ptr, nq, k, ef_search, tids_ptr, dist_ptr, cnt_ptr):
        """hb_search_batch_async on raw (pinned) host pointers; complete with search_wait(slot)."""
        self._ck(self._L.hb_search_batch_async(self._h, slot, C.c_void_p(q_ptr), nq, ef_search, k, C.c_void_p(tids_ptr),
                                               C.c_void_p(dist_ptr), C.c_void_p(cnt_ptr)), "hb_search_batch_async")

    def search_wait(self, slot):
        self._ck(self._L.hb_search_batch_wait(self._h, slot), "hb_search_batch_wait")

    def search_elements(self, queries, ef_search=40):
        q = self._vecs(queries)
        nq = q.shape[0]
        elem = np.empty((nq, ef_search), np.int32)
        dist = np.empty((nq, ef_search), np.float32)
        cnt = np.empty(nq, np.int32)
        self._ck(self._L.hb_search_batch_elements(self._h, _p(q), nq, ef_search, _p(elem), _p(dist), _p(cnt)),
                 "hb_search_batch_elements")
        return elem, dist, cnt

    def search_dev(self, dev_queries, nq, ef_search, dev_elem, dev_dist, dev_cnt, stream=0):
        """Device-resident batch: raw device pointers (ints), asynchronous on `stream`."""
        self._ck(self._L.hb_search_batch_dev(self._h, C.c_void_p(dev_queries), nq, ef_search, C.c_void_p(dev_elem),
                                             C.c_void_p(dev_dist), C.c_void_p(dev_cnt), C.c_void_p(stream)),
                 "hb_search_batch_dev")

    def search_status(self, stream=0):
        """waits for `stream`; raises if the last search_dev on it hit the tie-tail limit"""
        self._ck(self._L.hb_search_batch_status(self._h, C.c_void_p(stream)), "hb_search_batch_status")

    def elements_to_tids_dev(self, dev_elem, dev_dist, nq, ef, k, dev_tids, dev_tdist, stream=0):
        self._ck(self._L.hb_elements_to_tids_dev(self._h, C.c_void_p(dev_elem), C.c_void_p(dev_dist), nq, ef, k,
                                                 C.c_void_p(dev_tids), C.c_void_p(dev_tdist), C.c_void_p(stream)),
                 "hb_elements_to_tids_dev")

    def search_layer(self, queries, ep, ef, layer):
        q = self._vecs(queries)
        nq = q.shape[0]
        ep = np.ascontiguousarray(ep, np.int32).reshape(nq, -1)
        nep = ep.shape[1]
        stride = max(ef, nep)
        elem = np.empty((nq, stride), np.int32)
        dist = np.empty((nq, stride), np.float32)
        cnt = np.empty(nq, np.int32)
        self._ck(self._L.hb_search_layer(self._h, _p(q), nq, _p(ep), nep, ef, layer, _p(elem), _p(dist), _p(cnt)),
                 "hb_search_layer")
        return elem, dist, cnt

    def counters(self, reset=False):
        c = Counters()
        self._ck(self._L.hb_get_counters(self._h, C.byref(c), int(reset)), "hb_get_counters")
        return c.as_dict()

    def per_query_counters(self, nq):
        out = np.empty((nq, 4), np.int32)
        self._ck(self._L.hb_get_per_query_counters(self._h, nq, _p(out)), "hb_get_per_query_counters")
        return out

    def last_search_ms(self):
        return float(self._L.hb_last_search_ms(self._h))

    # ---- opclass support functions
    def distance(self, queries, cand):
        """FUNCTION 1 of the opclass, batched: queries x candidate-element lists -> distances."""
        q = self._vecs(queries)
        cand = np.ascontiguousarray(cand, np.int32).reshape(q.shape[0], -1)
        out = np.empty(cand.shape, np.float32)
        self._ck(self._L.hb_distance_batch(self._h, _p(q), q.shape[0], _p(cand), cand.shape[1], _p(out)), "hb_distance_batch")
        return out

    def distance_dev(self, dev_queries, nq, dev_cand, nc, dev_out, stream=0):
        self._ck(self._L.hb_distance_batch_dev(self._h, C.c_void_p(dev_queries), nq, C.c_void_p(dev_cand), nc,
                                               C.c_void_p(dev_out), C.c_void_p(stream)), "hb_distance_batch_dev")

    def normalize(self, vecs):
        """l2_normalize (FUNCTION 2 + normalisation) -> (normalised, ok mask)."""
        v = self._vecs(vecs)
        out = np.empty_like(v)
        ok = np.empty(v.shape[0], np.uint8)
        self._ck(self._L.hb_normalize(self._h, _p(v), v.shape[0], _p(out), _p(ok)), "hb_normalize")
        return out, ok.astype(bool)

    def bruteforce(self, queries, k=10, debug_scores=False, stats=False):
        """Exact scan of the partition (recall ground truth): bf16 tcgen05 GEMM candidates, fp32
        re-rank, certified by the bf16 error bound (uncertified queries are re-scanned in fp32)."""
        q = self._vecs(queries)
        nq = q.shape[0]
        elem = np.empty((nq, k), np.int32)
        dist = np.empty((nq, k), np.float32)
        dbg = np.empty((nq, self.n), np.float32) if debug_scores else None
        st = np.zeros(3, np.float32)
        self._ck(self._L.hb_bruteforce_ex(self._h, _p(q), nq, k, _p(elem), _p(dist), _p(dbg), _p(st)), "hb_bruteforce")
        out = (elem, dist)
        if debug_scores:
            out += (dbg,)
        if stats:
            out += ({"certified": int(st[0]), "rescanned": int(st[1]), "gemm_ms": float(st[2])},)
        return out


class HnswScan:
    """IndexScanDesc for one ordered scan: rescan binds the query, gettuple streams heap TIDs."""

    def __init__(self, index):
        self.index = index
        self._L = index._L
        self._h = self._L.hb_beginscan(index._h)
        if not self._h:
            raise _err(self._L, "hb_beginscan")

    def set_iterative(self, mode="relaxed_order", max_scan_tuples=20000):
        """hnsw.iterative_scan = off | relaxed_order | strict_order, hnsw.max_scan_tuples"""
        m = {"off": 0, "relaxed_order": 1, "strict_order": 2}[mode]
        if self._L.hb_scan_set_iterative(self._h, m, max_scan_tuples) < 0:
            raise _err(self._L, "hb_scan_set_iterative")

    def rescan(self, query, ef_search=40):
        q = self.index._vecs(query)
        if q.shape[0] != 1:
            raise HnswError("rescan takes one query vector")
        rc = self._L.hb_rescan(self._h, _p(q), ef_search)
        if rc < 0:
            raise _err(self._L, "hb_rescan")

    def gettuple(self):
        """-> (heap_tid, distance) or None when the scan is exhausted."""
        tid, d = C.c_int64(), C.c_float()
        rc = self._L.hb_gettuple(self._h, C.byref(tid), C.byref(d))
        if rc < 0:
            raise _err(self._L, "hb_gettuple")
        return (tid.value, d.value) if rc == 1 else None

    def endscan(self):
        if self._h:
            self._L.hb_endscan(self._h)
            self._h = None

    __del__ = endscan


def pgvector_pages_info(pages):
    """(dim, m, ef_construction, n_elements, upper_rows) of a pgvector HNSW index relation's pages; host only."""
    L = load_library()
    buf = np.frombuffer(pages, np.uint8) if isinstance(pages, (bytes, bytearray)) else np.ascontiguousarray(pages, np.uint8)
    if buf.size % 8192:
        raise HnswError("index pages must be a multiple of 8192 bytes")
    d, m, efc = C.c_int(), C.c_int(), C.c_int()
    ne, ur = C.c_int64(), C.c_int64()
    if L.hb_pgvector_pages_info(_p(buf), buf.size // 8192, C.byref(d), C.byref(m), C.byref(efc), C.byref(ne), C.byref(ur)) < 0:
        raise _err(L, "hb_pgvector_pages_info")
    return d.value, m.value, efc.value, ne.value, ur.value


class HnswIterator:
    """Batched resumable scans (hnsw.iterative_scan): next() returns (elem, dist, cnt), each query's
    next batch of elements nearest-first, or None when every scan is exhausted."""

    def __init__(self, index, queries, ef_search=40, max_scan_tuples=20000):
        self.index = index
        self._L = index._L
        q = index._vecs(queries)
        self.nq, self.ef = q.shape[0], ef_search
        self._h = self._L.hb_iter_begin(index._h, _p(q), self.nq, ef_search, max_scan_tuples)
        if not self._h:
            raise _err(self._L, "hb_iter_begin")

    def next(self):
        e = np.empty((self.nq, self.ef), np.int32)
        d = np.empty((self.nq, self.ef), np.float32)
        c = np.empty(self.nq, np.int32)
        got = self._L.hb_iter_next(self._h, _p(e), _p(d), _p(c))
        if got < 0:
            raise _err(self._L, "hb_iter_next")
        return None if got == 0 else (e, d, c)

    def tuples(self):
        t = np.empty(self.nq, np.int64)
        if self._L.hb_iter_tuples(self._h, _p(t)) < 0:
            raise _err(self._L, "hb_iter_tuples")
        return t

    def close(self):
        if self._h:
            self._L.hb_iter_end(self._h)
            self._h = None

    __del__ = close
