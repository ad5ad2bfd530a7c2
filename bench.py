#!/usr/bin/env python
"""bench.py -- headline benchmark of the HNSW hot path on B200.

Metric (BASELINE.json): QPS at recall@10 >= 0.95 on 1M x 768 fp32 cosine (configs[1]); also reports
HNSW build vectors/s.  One "step" = one pass of the batched scan (hnswgettuple for nq queries) over
one batch of synthetic queries.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun ... bench.py --gpus N ...        (one rank per GPU)

N = 1: the single 1M x 768 index of configs[1].  N > 1: a single unpartitioned graph does not
shard, so every rank holds a replica and scans its own query batches: "replicas
only", weak scaling, no data-path collective.  `--workload partitioned` runs the hash-partitioned
path instead (P partitions over the ranks, queries broadcast, one NCCL all-gather of per-rank top-k,
merge) -- the configs[2] shape.

`--impl reference`: the reference's CPU path.  The mount has no source and there is no PostgreSQL
(/root/reference/README.md:1), so what is timed is the C oracle (oracle/, "CPU restatement of
pgvector HNSW semantics -- not pgvector") with pgvector's natural summation order, on all host
threads, over a bounded sample of the same workload.  The 1M-element graph it searches is built by
the GPU builder during untimed set-up (a single-threaded CPU build of 1M x 768 takes hours); the
timed region runs no GPU code.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def gen_set(n, dim, seed, device, n_centres=4096, latent=64, spread=0.35, noise=0.1, centre_seed=20260101):
    """SURVEY.md 8(d) C2 generator: mixture of Gaussian centres in a low-dimensional latent,
    random projection to `dim`, isotropic noise.  Not normalised (the cosine opclass does that)."""
    import torch
    gc = torch.Generator(device=device).manual_seed(centre_seed)
    cent = torch.randn((n_centres, latent), generator=gc, device=device)
    proj = torch.randn((latent, dim), generator=gc, device=device) / latent ** 0.5
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n, dim), dtype=torch.float32, device=device)
    step = 1 << 18
    for s in range(0, n, step):
        e = min(n, s + step)
        a = torch.randint(0, n_centres, (e - s,), generator=g, device=device)
        z = cent[a] + spread * torch.randn((e - s, latent), generator=g, device=device)
        x = z @ proj
        x = x + noise * torch.randn((e - s, dim), generator=g, device=device)
        out[s:e] = x
    return out


def exact_topk(x_dev, q_dev, k):
    """fp32 exact cosine top-k with torch (checker for recall only)."""
    import torch
    xn = torch.nn.functional.normalize(x_dev, dim=1)
    qn = torch.nn.functional.normalize(q_dev, dim=1)
    out = []
    for s in range(0, qn.shape[0], 256):
        sims = qn[s:s + 256] @ xn.T
        out.append(torch.topk(sims, k, dim=1).indices)
    return torch.cat(out).cpu().numpy()


def recall_at(ids, gt):
    return float(np.mean([len(set(ids[i]) & set(gt[i])) / gt.shape[1] for i in range(gt.shape[0])]))


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_index(pkg, x_host, dim, device_index, opclass="vector_cosine_ops", m=16, efc=64, seed=1):
    # warm-up: a 4096-row throwaway index loads the build kernels' modules (CUDA loads them lazily at first
    # launch) so that the timed build measures the build
    warm = pkg.HnswIndex(dim, opclass, m, efc, capacity=4096, device=device_index, seed=seed)
    warm.build(x_host[:4096])
    warm.close()
    ix = pkg.HnswIndex(dim, opclass, m, efc, capacity=x_host.shape[0], device=device_index, seed=seed)
    t0 = time.time()
    n = ix.build(x_host)
    dt = time.time() - t0
    return ix, n, dt


def pick_ef(ix, q_dev, gt, nq_eval, efs, stream, torch, target=0.95):
    sweep = []
    chosen = None
    for ef in efs:
        elem = torch.empty((nq_eval, ef), dtype=torch.int32, device=q_dev.device)
        dist = torch.empty((nq_eval, ef), dtype=torch.float32, device=q_dev.device)
        cnt = torch.empty((nq_eval,), dtype=torch.int32, device=q_dev.device)
        ix.search_dev(q_dev.data_ptr(), nq_eval, ef, elem.data_ptr(), dist.data_ptr(), cnt.data_ptr(), stream)
        torch.cuda.synchronize()
        r = recall_at(elem[:, :10].cpu().numpy(), gt)
        sweep.append({"ef_search": ef, "recall@10": round(r, 4)})
        if chosen is None and r >= target:
            chosen = (ef, r)
            break
    if chosen is None:
        chosen = (efs[-1], sweep[-1]["recall@10"])
    return chosen[0], chosen[1], sweep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "partitioned", "build"])
    ap.add_argument("--rows", dest="n", type=int, default=1000000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--queries", dest="nq", type=int, default=10000)
    ap.add_argument("--ef", type=int, default=0, help="hnsw.ef_search (0 = smallest of the sweep reaching recall 0.95)")
    ap.add_argument("--partitions", type=int, default=8)
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (0 = sized for ~15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference" and rank != 0:
        return 0
    # stdout carries the one JSON line and nothing else: NCCL's banner ("NCCL version ...") goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

    import torch
    import pgvector_hnsw_partitioning_b200 as pkg

    if not torch.cuda.is_available():
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "no CUDA device to build the 1M-element graph the CPU path searches"}))
            return 0
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n, dim, nq, k = args.n, args.dim, args.nq, 10
    hbm_peak, peak_kind = peaks()
    base_seed = 20260101 + 1
    if args.workload == "partitioned" and args.impl == "ours":
        return run_partitioned(args, pkg, torch, dev, rank, local_rank, world, hbm_peak, peak_kind)
    if args.workload == "build" and args.impl == "ours":
        return run_build(args, pkg, torch, dev, rank, local_rank, world, hbm_peak, peak_kind)

    # ---------------------------------------------------------------- data + index (untimed)
    t0 = time.time()
    x_dev = gen_set(n, dim, base_seed, dev)
    x_host = x_dev.cpu().numpy()
    log("[rank %d] generated %d x %d in %.1fs" % (rank, n, dim, time.time() - t0))
    ix, n_indexed, build_s = build_index(pkg, x_host, dim, local_rank)
    bc = ix.counters(reset=True)
    log("[rank %d] built %d elements in %.1fs (%.0f vectors/s)" % (rank, n_indexed, build_s, n_indexed / build_s))
    row_bytes = dim * 4
    build_bytes = (bc["n_dist"] + bc["n_pair"]) * row_bytes

    # queries: a distinct batch per step, different for every rank
    total_steps = args.warmup + args.steps
    nq_eval = min(4000, nq)      # recall is estimated on 4000 queries: the 0.95 threshold is decided within ~0.2 pt
    q_eval = gen_set(nq_eval, dim, base_seed + 1000, dev)
    # ground truth: the library's exact scan (bf16 tcgen05 GEMM + fp32 re-rank, certified), cross-checked
    # against a plain torch fp32 scan
    gt_torch = exact_topk(x_dev, q_eval, k)
    del x_dev
    gt, _, bf_stats = ix.bruteforce(q_eval.cpu().numpy(), k, stats=True)
    gt_agree = recall_at(gt, gt_torch)
    exact = None
    if args.impl == "ours" and rank == 0:
        qbf = gen_set(nq, dim, base_seed + 3000, dev).cpu().numpy()
        ix.bruteforce(qbf, k)
        t0 = time.perf_counter()
        _, _, st = ix.bruteforce(qbf, k, stats=True)
        bf_s = time.perf_counter() - t0
        tf = 2.0 * nq * n * dim / (st["gemm_ms"] * 1e-3) / 1e12
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                tpeak = float(json.load(f)["bf16_tflops_sustained"])
        except Exception:
            tpeak = 1400.0
        exact = {"queries_per_s": round(nq / bf_s, 1), "gemm_ms": round(st["gemm_ms"], 3), "gemm_tflops": round(tf, 1),
                 "tensor_peak_tflops_sustained": tpeak, "frac_of_tensor_peak": round(tf / tpeak, 4),
                 "certified_exact": st["certified"], "rescanned_fp32": st["rescanned"], "batch": nq,
                 "agreement_with_torch_fp32_top10": round(gt_agree, 5)}
    torch.cuda.empty_cache()
    stream = torch.cuda.current_stream().cuda_stream
    efs = [args.ef] if args.ef > 0 else [40, 50, 60, 70, 80, 90, 100, 120, 150, 200, 300, 400]
    ef, rec, sweep = pick_ef(ix, q_eval, gt, nq_eval, efs, stream, torch)
    log("[rank %d] ef_search=%d recall@10=%.4f sweep=%s" % (rank, ef, rec, sweep))

    if args.impl == "reference":
        return run_reference(args, pkg, ix, x_host, q_eval, ef, rec, n, dim, nq, build_s)

    del x_host
    q_all = gen_set(nq * total_steps, dim, base_seed + 2000 + rank, dev).view(total_steps, nq, dim)
    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident steps
    # Steps alternate between two CUDA streams (each with its own output buffers), the way a server
    # keeps two batches in flight: the drain of one batch overlaps the ramp of the next.
    NSTREAM = 3
    streams = [torch.cuda.Stream(device=dev) for _ in range(NSTREAM)]
    outs = [(torch.empty((nq, ef), dtype=torch.int32, device=dev), torch.empty((nq, ef), dtype=torch.float32, device=dev),
             torch.empty((nq,), dtype=torch.int32, device=dev)) for _ in range(NSTREAM)]
    main = torch.cuda.current_stream(dev)

    def run_steps(first, count):
        for st in streams:
            st.wait_stream(main)
        for s in range(count):
            st, (e_, d_, c_) = streams[s % NSTREAM], outs[s % NSTREAM]
            ix.search_dev(q_all[first + s].data_ptr(), nq, ef, e_.data_ptr(), d_.data_ptr(), c_.data_ptr(), st.cuda_stream)
        for st in streams:
            main.wait_stream(st)

    run_steps(0, args.warmup)
    torch.cuda.synchronize()
    ix.counters(reset=True)
    clocks = ClockSampler(local_rank)
    time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(main)
    run_steps(args.warmup, args.steps)
    ev1.record(main)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    # duration of the scan kernels of the last launch on its own stream (library events)
    last_kernel_ms = ix.last_search_ms()
    ctr = ix.counters(reset=True)
    clk = clocks.stop()
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        total_ms = float(t.item())
    qps = world * nq * args.steps / (total_ms / 1e3)

    # algorithmic bytes: distance evaluations x row + neighbour lists + the query
    alg_bytes = (ctr["n_dist"] * row_bytes + ctr["n_hop0"] * (2 * 16 * 4) + ctr["n_hopu"] * (16 * 4) + nq * args.steps * row_bytes)
    alg_per_launch = alg_bytes / args.steps
    ms_per_step = total_ms / args.steps
    achieved = alg_per_launch / (ms_per_step / 1e3) / 1e9

    # ---------------------------------------------------------------- end to end through the C ABI
    # host (pinned) buffers in and out; two batches in flight (hb_search_batch_async slots 0/1), so
    # every step's H2D copy, scan and D2H read are inside the timed region and overlap one another.
    qh = torch.empty((total_steps, nq, dim), dtype=torch.float32).pin_memory()
    qh.copy_(q_all.cpu())
    houts = [(torch.empty((nq, k), dtype=torch.int64).pin_memory(), torch.empty((nq, k), dtype=torch.float32).pin_memory(),
              torch.empty((nq,), dtype=torch.int32).pin_memory()) for _ in range(NSTREAM)]

    def run_e2e(first, count):
        for s in range(count):
            slot = s % NSTREAM
            ix.search_wait(slot)
            t_, d_, c_ = houts[slot]
            ix.search_async(slot, qh[first + s].data_ptr(), nq, k, ef, t_.data_ptr(), d_.data_ptr(), c_.data_ptr())
        for slot in range(NSTREAM):
            ix.search_wait(slot)

    run_e2e(0, args.warmup)
    barrier()
    t0 = time.perf_counter()
    run_e2e(args.warmup, args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_qps = world * nq * args.steps / e2e_s
    e2e_recall = None
    if rank == 0:
        out_t, out_d, out_c = houts[0]
        ix.search_into(q_eval.cpu().pin_memory().data_ptr(), nq_eval, k, ef, out_t.data_ptr(), out_d.data_ptr(), out_c.data_ptr())
        e2e_recall = recall_at(out_t[:nq_eval].numpy(), gt)

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(ix, q_eval.cpu().numpy(), ef, args.cpu_sample)
        except Exception as e:   # the baseline is reported, never required
            cpu = {"error": str(e)}

    if rank == 0:
        line = {
            "metric": "QPS @ recall@10>=0.95 (1M x 768 cosine)", "value": round(qps, 1), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: %dx%d fp32 cosine, m=16, ef_construction=64, ef_search=%d, k=10, batch=%d "
                                   "queries/step resident in HBM, steps rotate over 3 streams%s" % (n, dim, ef, nq, "" if world == 1 else ", one replica per GPU (replicas only)"),
                       "ef_search": ef, "recall@10": round(rec, 4), "recall_sweep": sweep, "parallelism": "replicas x%d" % world,
                       "l2_policy": "inputs larger than L2: graph+vectors %.2f GB, a distinct query batch every step" % ((n * row_bytes + n * 128) / 1e9),
                       "parity": "unpinned (reference mount has no source); ids bit-identical to oracle/ in tests"},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                         "frac": round(achieved / hbm_peak, 4),
                         # dram__bytes_read.sum + dram__bytes_write.sum of one scan_kernel launch of this workload
                         # (10 000 queries) from the ncu captures profiles/r1_v3 (ef_search=100) and r1_v4 (ef_search=90)
                         "traffic": ({100: 13080542432, 90: 12828498888}.get(ef) if (n, dim, nq) == (1000000, 768, 10000) else None),
                         "peak_kind": peak_kind,
                         "frac_of_nominal_8000": round(achieved / 8000.0, 4),
                         "algorithmic_bytes_per_launch": int(alg_per_launch), "kernel": "scan_kernel (batched HnswSearchLayer)",
                         "last_launch_ms": round(last_kernel_ms, 4),
                         "n_dist_per_query": round(ctr["n_dist"] / (nq * args.steps), 1),
                         "n_hop_per_query": round((ctr["n_hop0"] + ctr["n_hopu"]) / (nq * args.steps), 1),
                         "slow_path_queries": ctr["n_slow"]},
            "cpu_baseline": cpu,
            "e2e": {"value": round(e2e_qps, 1), "unit": "queries/s", "h2d_bytes_per_step": nq * row_bytes,
                    "d2h_bytes_per_step": nq * (k * 12 + 4), "recall@10": e2e_recall},
            "exact_scan": exact,
            "gpu_launches": 3 * args.steps,
            "clocks": clk,
            "build": {"vectors_per_s": round(n_indexed / build_s, 1), "seconds": round(build_s, 2), "n": n_indexed,
                      "algorithmic_gb": round(build_bytes / 1e9, 1), "achieved_gbs": round(build_bytes / build_s / 1e9, 1)},
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def oracle_from_index(ix, mode):
    from oracle import oracle as O
    g = ix.export_graph()
    return O, O.Index.from_graph(g, mode)


def cpu_baseline(ix, q_eval, ef, sample):
    """The oracle (kind "port") on the host cores over a bounded sample of the same workload."""
    cores = os.cpu_count() or 1
    O, orc = oracle_from_index(ix, 1)   # NATURAL: pgvector's scalar loops with its compiler flags
    probe = q_eval[:min(64, len(q_eval))]
    t0 = time.perf_counter()
    orc.search_batch(probe, ef, threads=cores)
    per_q = (time.perf_counter() - t0) / len(probe)
    if sample <= 0:
        sample = int(max(200, min(len(q_eval), 15.0 / max(per_q, 1e-6))))
    qs = q_eval[:sample]
    t0 = time.perf_counter()
    orc.search_batch(qs, ef, threads=cores)
    dt = time.perf_counter() - t0
    one = qs[:min(len(qs), max(50, int(3.0 / max(per_q * cores, 1e-6))))]
    t0 = time.perf_counter()
    orc.search_batch(one, ef, threads=1)
    dt1 = time.perf_counter() - t0
    return {"value": round(len(qs) / dt, 1), "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": "%d queries of the same distribution, ef_search=%d, all %d host threads (one backend per thread); "
                      "single thread: %.1f queries/s. CPU restatement of pgvector HNSW semantics, not pgvector: no buffer "
                      "manager/WAL, so faster than the real extension" % (len(qs), ef, cores, len(one) / dt1),
            "single_thread_value": round(len(one) / dt1, 1)}


def run_reference(args, pkg, ix, x_host, q_eval, ef, rec, n, dim, nq, build_s):
    """--impl reference: the CPU oracle timed over bounded samples, all host threads."""
    import torch
    cores = os.cpu_count() or 1
    O, orc = oracle_from_index(ix, 1)
    ix.close()
    torch.cuda.empty_cache()
    qs_all = q_eval.cpu().numpy()
    probe = qs_all[:64]
    t0 = time.perf_counter()
    orc.search_batch(probe, ef, threads=cores)
    per_q = (time.perf_counter() - t0) / len(probe)
    steps_total = args.warmup + args.steps
    # bounded: the whole run ~20-40 s of CPU work
    per_step = int(max(64, min(len(qs_all), 30.0 / max(per_q, 1e-6) / max(steps_total, 1))))
    times = []
    for s in range(steps_total):
        lo = (s * per_step) % max(1, len(qs_all) - per_step + 1)
        t0 = time.perf_counter()
        orc.search_batch(qs_all[lo:lo + per_step], ef, threads=cores)
        times.append(time.perf_counter() - t0)
    timed = times[args.warmup:]
    qps = per_step * len(timed) / sum(timed)
    sample = "%d queries per step (bounded sample of the %d-query batch), ef_search=%d, %d host threads" % (per_step, nq, ef, cores)
    line = {"impl": "reference", "metric": "QPS @ recall@10>=0.95 (1M x 768 cosine)", "value": round(qps, 1), "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * sum(timed) / len(timed), 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: %dx%d fp32 cosine, m=16, ef_construction=64, ef_search=%d, k=10" % (n, dim, ef),
                       "ef_search": ef, "recall@10": round(rec, 4),
                       "note": "CPU restatement of pgvector HNSW semantics (oracle/), not pgvector: the reference mount has no "
                               "source and the image has no PostgreSQL. Graph built by the GPU builder in untimed set-up."},
            "cpu_baseline": {"value": round(qps, 1), "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(qps, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def exact_topk_metric(x_dev, q_dev, k, metric):
    """torch fp32 exact top-k for l2 / ip / cosine (checker for recall only)."""
    import torch
    x = x_dev.float()
    q = q_dev.float()
    if metric == "cosine":
        x = torch.nn.functional.normalize(x, dim=1)
        q = torch.nn.functional.normalize(q, dim=1)
    xx = (x * x).sum(1) if metric == "l2" else None
    out = []
    for s in range(0, q.shape[0], 256):
        sims = q[s:s + 256] @ x.T
        if metric == "l2":
            sims = 2 * sims - xx[None, :]
        out.append(torch.topk(sims, k, dim=1).indices)
    return torch.cat(out).cpu().numpy()


def run_build(args, pkg, torch, dev, rank, local_rank, world, hbm_peak, peak_kind):
    """configs[3] shape: HNSW index build, 1M x 1536 halfvec inner product by default, hash-partitioned
    into --partitions partitions that the ranks build independently (no collective); then a merged
    search at ef_search=40 checks recall of what was built.  `value` = rows / max-over-ranks build time."""
    n, k, P = args.n, 10, args.partitions
    dim = args.dim if args.dim != 768 else 1536
    opclass = os.environ.get("HB_BUILD_OPCLASS", "halfvec_ip_ops")
    half = opclass.startswith("halfvec")
    metric = "ip" if "_ip_" in opclass else ("l2" if "_l2_" in opclass else "cosine")
    x = gen_set(n, dim, 20260104, dev)
    if metric == "ip":     # not normalised: norms ~ lognormal(sigma = 0.1)
        x = x * torch.exp(0.1 * torch.randn((n, 1), device=dev, generator=torch.Generator(device=dev).manual_seed(5)))
    xs = x.half() if half else x
    x_host = xs.cpu().numpy()
    q = gen_set(1000, dim, 20260104 + 1000, dev)
    qs = q.half() if half else q
    gt = exact_topk_metric(xs, qs, k, metric)
    del x, xs
    torch.cuda.empty_cache()
    pix = pkg.PartitionedIndex(dim, opclass, P, 16, 64, capacity_per_partition=int(n / P * 1.1) + 1024, rank=rank, world=world,
                               device=local_rank, seed=3)
    warm = pkg.HnswIndex(dim, opclass, 16, 64, capacity=4096, device=local_rank, seed=3)     # loads the kernels' modules
    warm.build(x_host[:4096])
    warm.close()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pix.build(x_host)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    ctr = {"n_dist": 0, "n_pair": 0}
    for ix in pix.parts.values():
        c = ix.counters(reset=True)
        ctr["n_dist"] += c["n_dist"]; ctr["n_pair"] += c["n_pair"]
    tt = torch.tensor([build_s, float(ctr["n_dist"]), float(ctr["n_pair"])], device=dev, dtype=torch.float64)
    if world > 1:
        tmax = tt.clone()
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.SUM)
        build_s = float(tmax[0].item())
    n_dist, n_pair = float(tt[1].item()), float(tt[2].item())
    t, d = pix.search_dev(qs, k, 40)
    rec = recall_at(t.cpu().numpy(), gt)
    row_bytes = dim * (2 if half else 4)
    alg = (n_dist + n_pair) * row_bytes
    if rank == 0:
        print(json.dumps({"metric": "HNSW build vectors/s", "value": round(n / build_s, 1), "unit": "vectors/s", "n_gpus": world,
                          "steps": 1, "warmup": 0, "ms_per_step": round(build_s * 1e3, 1), "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f16" if half else "f32", "data": "synthetic",
                          "config": {"workload": "configs[3]: %dx%d %s index build, m=16, ef_construction=64, %d hash partitions "
                                                 "built independently by %d rank(s), host rows in, no collective" % (n, dim, opclass, P, world),
                                     "parallelism": "partitions/%d" % world, "recall@10_ef40_merged": round(rec, 4)},
                          "roofline": {"bound": "hbm", "achieved": round(alg / build_s / 1e9 / world, 1), "peak": hbm_peak, "unit": "GB/s",
                                       "frac": round(alg / build_s / 1e9 / world / hbm_peak, 4), "traffic": None, "peak_kind": peak_kind,
                                       "note": "algorithmic bytes = (n_dist + n_pair) x row, the sequential algorithm's evaluations "
                                               "(oracle counters), per GPU; the link phase memoises pair distances, so fewer rows are "
                                               "actually fetched and the figure can exceed the HBM peak",
                                       "n_dist_per_insert": round(n_dist / n, 1), "n_pair_per_insert": round(n_pair / n, 1)},
                          "gpu_launches": None}))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def run_partitioned(args, pkg, torch, dev, rank, local_rank, world, hbm_peak, peak_kind):
    """configs[2] shape: P hash partitions over the ranks, queries broadcast, NCCL all-gather of the
    per-rank top-k, merge.  Default sizes are scaled by --n (total rows)."""
    n, dim, nq, k, P = args.n, args.dim, args.nq, 10, args.partitions
    ef = args.ef if args.ef > 0 else 40
    pix = pkg.PartitionedIndex(dim, "vector_l2_ops" if dim == 128 else "vector_cosine_ops", P, 16, 64,
                               capacity_per_partition=int(n / P * 1.1) + 1024, rank=rank, world=world, device=local_rank, seed=3)
    x_dev = gen_set(n, dim, 20260103, dev)
    q_eval = gen_set(1000, dim, 20260103 + 500, dev)
    gt = exact_topk_metric(x_dev, q_eval, k, "l2" if dim == 128 else "cosine")
    x = x_dev.cpu().numpy()
    del x_dev
    t0 = time.time()
    pix.build(x)
    build_s = time.time() - t0
    te, _ = pix.search_dev(q_eval, k, ef)
    rec = recall_at(te.cpu().numpy(), gt)
    for ix in pix.parts.values():
        ix.counters(reset=True)
    total_steps = args.warmup + args.steps
    q_all = gen_set(nq * total_steps, dim, 20260103 + 1000, dev).view(total_steps, nq, dim)   # same on every rank = broadcast
    for w in range(args.warmup):
        pix.search_dev(q_all[w], k, ef)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(args.steps):
        t, d = pix.search_dev(q_all[args.warmup + s], k, ef)
    ev1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        ms = float(tt.item())
    cs = [ix.counters(reset=True) for ix in pix.parts.values()]
    row_bytes = dim * 4
    alg = sum(c["n_dist"] * row_bytes + c["n_hop0"] * 128 + c["n_hopu"] * 64 for c in cs) + len(cs) * nq * args.steps * row_bytes
    at = torch.tensor([float(alg)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(at, op=torch.distributed.ReduceOp.SUM)
    alg_gbs_per_gpu = float(at.item()) / (ms / 1e3) / 1e9 / world
    if rank == 0:
        print(json.dumps({"metric": "QPS (hash-partitioned, %d partitions, merged top-%d)" % (P, k), "value": round(nq * args.steps / (ms / 1e3), 1),
                          "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "partitioned: %dx%d in %d hash partitions, ef_search=%d, queries broadcast, "
                                                 "all-gather + merge" % (n, dim, P, ef), "parallelism": "partitions/%d" % world,
                                     "recall@10": round(rec, 4)},
                          "roofline": {"bound": "hbm", "achieved": round(alg_gbs_per_gpu, 1), "peak": hbm_peak, "unit": "GB/s",
                                       "frac": round(alg_gbs_per_gpu / hbm_peak, 4), "traffic": None, "peak_kind": peak_kind,
                                       "note": "per GPU; every query visits every partition, so the step does P searches per query"},
                          "build": {"seconds": round(build_s, 2), "vectors_per_s": round(n / build_s, 1)},
                          "gpu_launches": args.steps * (len(pix.owned) * 4 + 2)}))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
