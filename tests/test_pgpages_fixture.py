"""A hand-assembled pgvector HNSW index relation (two 8 kB blocks), written out field by field from the struct offsets
documented in csrc/pgpages.cu -- PageHeaderData, ItemIdData, HnswPageOpaqueData, HnswMetaPageData,
HnswElementTupleData, HnswNeighborTupleData, ItemPointerData -- independently of tests/pgpages_writer.py.  The bytes
are also committed as tests/golden/pgvector_pages_tiny.bin.  Still a RECALLED layout (no PostgreSQL in the image): it
pins the reader to the documented offsets, not to pgvector."""
import os
import struct

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BLCKSZ = 8192
DIM, M, EFC = 4, 2, 8
# three elements: vectors, levels, heap TIDs (block, offset), neighbour lists per layer (element numbers)
VECS = [(1.0, 0.0, 0.0, 0.0), (0.0, 1.0, 0.0, 0.0), (0.0, 0.0, 2.0, 0.0)]
LEVELS = [1, 0, 0]
HEAPTIDS = [[(7, 1), (7, 9)], [(7, 2)], [(8, 1)]]
NBRS = {0: {1: [], 0: [1, 2]}, 1: {0: [0, 2]}, 2: {0: [1, 0]}}
# where things sit on block 1: element tuples at items 1, 3, 5, their neighbour tuples at items 2, 4, 6
ELEM_ITEM = {0: 1, 1: 3, 2: 5}


def tid(blk, off):
    return struct.pack("<HHH", blk >> 16, blk & 0xffff, off)


def element_tuple(e):
    t = struct.pack("<BBBB", 1, LEVELS[e], 0, 1)                               # type, level, deleted, version
    tids = HEAPTIDS[e] + [(0, 0)] * (10 - len(HEAPTIDS[e]))
    t += b"".join(tid(b, o) for b, o in tids)                                  # heaptids[10] at +4
    t += tid(1, ELEM_ITEM[e] + 1)                                              # neighbortid at +64
    t += struct.pack("<H", 0)                                                  # unused at +70
    t += struct.pack("<ihh", (8 + 4 * DIM) << 2, DIM, 0) + struct.pack("<%df" % DIM, *VECS[e])   # vector datum at +72
    return t


def neighbor_tuple(e):
    slots = []
    for lc in range(LEVELS[e], -1, -1):
        lm = 2 * M if lc == 0 else M
        ids = NBRS[e][lc]
        slots += [tid(1, ELEM_ITEM[x]) for x in ids] + [tid(0, 0)] * (lm - len(ids))
    assert len(slots) == (LEVELS[e] + 2) * M
    return struct.pack("<BBH", 2, 1, len(slots)) + b"".join(slots)             # type, version, count, indextids[]


def page(tuples):
    buf = bytearray(BLCKSZ)
    special = BLCKSZ - 8
    upper = special
    items = b""
    for t in tuples:
        ln = len(t)
        upper -= (ln + 7) & ~7                                                 # MAXALIGN
        buf[upper:upper + ln] = t
        items += struct.pack("<I", upper | (1 << 15) | (ln << 17))             # lp_off:15 | lp_flags:2 = LP_NORMAL | lp_len:15
    lower = 24 + len(items)
    buf[24:lower] = items
    struct.pack_into("<HHH", buf, 12, lower, upper, special)                   # pd_lower, pd_upper, pd_special
    struct.pack_into("<IHH", buf, special, 0xFFFFFFFF, 0, 0xFF90)              # HnswPageOpaqueData
    return bytes(buf)


def build_pages():
    meta = bytearray(BLCKSZ)
    struct.pack_into("<HHH", meta, 12, 24 + 28, BLCKSZ - 8, BLCKSZ - 8)
    # HnswMetaPageData at +24: magic, version, dimensions, m, efConstruction, entryBlkno, entryOffno, entryLevel, insertPage
    struct.pack_into("<IIIHHIHhI", meta, 24, 0xA953A953, 1, DIM, M, EFC, 1, ELEM_ITEM[0], LEVELS[0], 1)
    struct.pack_into("<IHH", meta, BLCKSZ - 8, 0xFFFFFFFF, 0, 0xFF90)
    tuples = []
    for e in range(3):
        tuples += [element_tuple(e), neighbor_tuple(e)]
    return bytes(meta) + page(tuples)


def test_fixture_bytes_are_the_committed_ones_and_info_parses(pkg):
    blob = build_pages()
    path = os.path.join(HERE, "golden", "pgvector_pages_tiny.bin")
    if not os.path.exists(path):                 # first run writes the fixture; it is committed
        open(path, "wb").write(blob)
    assert open(path, "rb").read() == blob
    assert pkg.pgvector_pages_info(blob) == (DIM, M, EFC, 3, 1)


@pytest.mark.gpu
def test_fixture_loads_and_searches(pkg):
    blob = open(os.path.join(HERE, "golden", "pgvector_pages_tiny.bin"), "rb").read()
    ix = pkg.HnswIndex(DIM, "vector_l2_ops", M, EFC, capacity=4)
    ix.load_pgvector_pages(blob)
    g = ix.export_graph()
    assert (g.n, g.entry, g.entry_level, g.upper_rows) == (3, 0, 1, 1)
    assert g.level.tolist() == LEVELS and g.ntids.tolist() == [2, 1, 1]
    assert g.nbr0.tolist() == [[1, 2, -1, -1], [0, 2, -1, -1], [1, 0, -1, -1]]
    assert g.uoff.tolist() == [0, -1, -1] and g.nbru[0].tolist() == [-1, -1]
    assert (g.vecs == np.array(VECS, np.float32)).all()
    assert g.tids[0, :2].tolist() == [(7 << 16) | 1, (7 << 16) | 9] and g.tids[2, 0] == (8 << 16) | 1
    t, d, c = ix.search(np.array([[0.0, 0.9, 0.0, 0.0]], np.float32), 4, 8)
    # nearest first; the element with two heap TIDs emits the later one first (hnswgettuple)
    assert t[0, :c[0]].tolist() == [(7 << 16) | 2, (7 << 16) | 9, (7 << 16) | 1, (8 << 16) | 1]
    # a halfvec handle must refuse these pages (element tuples too short / wrong type), not read garbage
    hx = pkg.HnswIndex(DIM * 4, "vector_l2_ops", M, EFC, capacity=4)
    with pytest.raises(pkg.HnswError):
        hx.load_pgvector_pages(blob)
    ix.close(); hx.close()
