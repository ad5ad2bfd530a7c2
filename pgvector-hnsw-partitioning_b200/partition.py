"""Hash-partitioned HNSW: the fork's partition routing and fan-out/merge for the hot path.

The reference mount has no source (/root/reference/README.md:1), so the partitioning contract is
the one BASELINE.json states: a row goes to partition splitmix64(heap_tid) mod P; a query is
broadcast to every partition; per-partition top-k lists are merged.  One process per GPU: rank r
owns partitions {p : p mod world == r}; the only collective on the data path is one all-gather of
nq x k (tid, distance) pairs per rank (NCCL over NVLink when the tensors are CUDA tensors).

torch is used for device buffers, streams and torch.distributed only.
"""
import numpy as np

from .hnsw import HB_F32, OPCLASSES, HnswError, HnswIndex, merge_topk_dev, partition_route


def owned_partitions(n_partitions, rank, world):
    return [p for p in range(n_partitions) if p % world == rank]


def exchange_topk(local_tids, local_dist, world, group=None):
    """The one data-path collective: all-gather of every rank's nq x k (tid, distance) list into
    world x nq x k.  Works on CUDA tensors (NCCL) and, for the host-logic tests, CPU tensors (gloo)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_tids[None], local_dist[None]
    nq, k = local_tids.shape
    all_t = torch.empty((world, nq, k), dtype=local_tids.dtype, device=local_tids.device)
    all_d = torch.empty((world, nq, k), dtype=local_dist.dtype, device=local_dist.device)
    dist.all_gather_into_tensor(all_t.view(world * nq, k), local_tids.contiguous(), group=group)
    dist.all_gather_into_tensor(all_d.view(world * nq, k), local_dist.contiguous(), group=group)
    return all_t, all_d


def split_rows(heap_tids, n_partitions, rank, world):
    """Rows this rank indexes: {partition -> row indices}, for the partitions it owns."""
    part = partition_route(heap_tids, n_partitions)
    return {p: np.nonzero(part == p)[0] for p in owned_partitions(n_partitions, rank, world)}


class PartitionedIndex:
    def __init__(self, dim, opclass="vector_l2_ops", n_partitions=8, m=16, ef_construction=64,
                 capacity_per_partition=1 << 20, rank=0, world=1, device=0, seed=0, group=None):
        if n_partitions < 1 or n_partitions > 64:
            raise HnswError("n_partitions must be in [1, 64]")
        self.dim, self.opclass, self.P, self.rank, self.world, self.device = dim, opclass, n_partitions, rank, world, device
        self.group = group
        self.metric, self.dtype = OPCLASSES[opclass]
        self.owned = owned_partitions(n_partitions, rank, world)
        self.parts = {p: HnswIndex(dim, opclass, m, ef_construction, capacity_per_partition, device, seed + p)
                      for p in self.owned}

    def close(self):
        for ix in self.parts.values():
            ix.close()
        self.parts = {}

    # ---- build: route rows to partitions, each rank indexes the partitions it owns; no collective
    def build(self, vecs, heap_tids=None):
        """The partitions a rank owns are built concurrently (one host thread per partition: the handles
        are independent and each has its own streams), so that the latency-bound small batches at the
        start of every build overlap on the GPU."""
        from concurrent.futures import ThreadPoolExecutor
        n = vecs.shape[0]
        tids = np.arange(n, dtype=np.int64) if heap_tids is None else np.ascontiguousarray(heap_tids, np.int64)
        rows = split_rows(tids, self.P, self.rank, self.world)

        def one(p):
            sel = rows[p]
            return self.parts[p].insert(vecs[sel], tids[sel]) if len(sel) else 0

        if len(self.parts) <= 1:
            return sum(one(p) for p in self.parts)
        with ThreadPoolExecutor(max_workers=min(len(self.parts), 8)) as ex:
            return sum(ex.map(one, list(self.parts)))

    @property
    def n_local(self):
        return sum(ix.n for ix in self.parts.values())

    # ---- search
    def search_local_dev(self, q_dev, k, ef_search):
        """q_dev: CUDA tensor nq x dim of the index dtype on this rank's GPU.  Returns this rank's
        merged (tids, dist) CUDA tensors, nq x k.  The owned partitions are searched on up to four
        CUDA streams at once (the drain of one partition's scan overlaps the ramp of the next), then
        merged on the caller's stream."""
        import torch
        nq = q_dev.shape[0]
        dev = q_dev.device
        main = torch.cuda.current_stream(dev)
        npart = len(self.parts)
        key = (nq, k, ef_search, str(dev))
        if getattr(self, "_buf_key", None) != key:
            n1 = max(npart, 1)
            self._bufs = (torch.empty((n1, nq, k), dtype=torch.int64, device=dev), torch.empty((n1, nq, k), dtype=torch.float32, device=dev),
                          torch.empty((n1, nq, ef_search), dtype=torch.int32, device=dev),
                          torch.empty((n1, nq, ef_search), dtype=torch.float32, device=dev),
                          torch.empty((n1, nq), dtype=torch.int32, device=dev),
                          torch.empty((nq, k), dtype=torch.int64, device=dev), torch.empty((nq, k), dtype=torch.float32, device=dev))
            self._buf_key = key
        if not getattr(self, "_streams", None):
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
        tids, dist, elem, edist, cnt, out_t, out_d = self._bufs
        tids.fill_(-1)
        dist.fill_(float("inf"))
        for st in self._streams:
            st.wait_stream(main)
        # work items = (partition, slice of the query batch).  Cutting the batch when a rank owns few
        # partitions was measured at 8 GPUs (one partition each) and did not pay: one slice per partition.
        nchunk = 1
        bounds = [nq * c // nchunk for c in range(nchunk + 1)]
        esz = q_dev.element_size() * q_dev.shape[1]
        item = 0
        for i, ix in enumerate(self.parts.values()):
            if ix.n == 0:
                continue
            for c in range(nchunk):
                lo, hi = bounds[c], bounds[c + 1]
                if hi <= lo:
                    continue
                st = self._streams[item % len(self._streams)].cuda_stream
                item += 1
                ix.search_dev(q_dev.data_ptr() + lo * esz, hi - lo, ef_search, elem[i, lo].data_ptr(), edist[i, lo].data_ptr(),
                              cnt[i, lo:].data_ptr(), st)
                ix.elements_to_tids_dev(elem[i, lo].data_ptr(), edist[i, lo].data_ptr(), hi - lo, ef_search, k,
                                        tids[i, lo].data_ptr(), dist[i, lo].data_ptr(), st)
        for st in self._streams:
            main.wait_stream(st)
        if npart <= 1:
            return tids[0], dist[0]
        merge_topk_dev(self.device, tids.data_ptr(), dist.data_ptr(), npart, nq, k, out_t.data_ptr(), out_d.data_ptr(),
                       main.cuda_stream)
        return out_t, out_d

    def exchange(self, local_tids, local_dist):
        return exchange_topk(local_tids, local_dist, self.world, self.group)

    def search_dev(self, q_dev, k=10, ef_search=40):
        """Broadcast queries are assumed resident on every rank.  Returns merged nq x k CUDA tensors
        (identical on every rank)."""
        import torch
        lt, ld = self.search_local_dev(q_dev, k, ef_search)
        all_t, all_d = self.exchange(lt, ld)
        if self.world == 1:
            return lt, ld
        nq = q_dev.shape[0]
        out_t = torch.empty((nq, k), dtype=torch.int64, device=q_dev.device)
        out_d = torch.empty((nq, k), dtype=torch.float32, device=q_dev.device)
        stream = torch.cuda.current_stream(q_dev.device).cuda_stream
        merge_topk_dev(self.device, all_t.data_ptr(), all_d.data_ptr(), self.world, nq, k, out_t.data_ptr(),
                       out_d.data_ptr(), stream)
        return out_t, out_d

    def search(self, queries, k=10, ef_search=40):
        """Host arrays in/out (H2D + D2H inside)."""
        import torch
        tdt = torch.float32 if self.dtype == HB_F32 else torch.float16
        q = torch.as_tensor(np.ascontiguousarray(queries)).to(tdt)
        q_dev = q.to("cuda:%d" % self.device, non_blocking=True)
        t, d = self.search_dev(q_dev, k, ef_search)
        return t.cpu().numpy(), d.cpu().numpy()
