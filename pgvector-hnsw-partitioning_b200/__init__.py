"""B200-native HNSW hot path behind pgvector's operator-class / access-method surface.

This package is a thin ctypes mirror of the C ABI in include/hnsw_b200.h (libhnsw_b200.so, built
from csrc/ for sm_100a).  Names follow the PostgreSQL index-AM callbacks the reference extension
implements (hnswbuild / hnswinsert / hnswbeginscan / hnswrescan / hnswgettuple / hnswendscan) and
pgvector's operator classes (vector_l2_ops, vector_ip_ops, vector_cosine_ops, halfvec_*).

There is no CPU fallback: importing works anywhere (so the symbol table can be checked), but any
compute call without the shared library or without a CUDA device raises.
"""
from .hnsw import (  # noqa: F401
    HB_COSINE, HB_F16, HB_F32, HB_HEAPTIDS, HB_IP, HB_L1, HB_L2, OPCLASSES, HnswError, HnswIndex, HnswIterator, HnswScan,
    build_library, pgvector_pages_info, lib_path, load_library, partition_of, partition_route, merge_topk_dev,
)
from .partition import PartitionedIndex, merge_rule, owned_partitions, share_unique_id, split_rows  # noqa: F401
