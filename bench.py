#!/usr/bin/env python
"""bench.py -- headline benchmark of the hilbert-quantization hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--rows R --dim D --queries Q] [--bf16-only] [--skip-map-index] [--skip-latency]

Default workload (BASELINE.json configs[1], "C2"): 1 M synthetic 1536-D embeddings mapped to 64x64
Hilbert grids with variant-C hierarchical indices, batches of 1024 queries, progressive
top-10 (coarse index filter -> cosine rerank -> top-k).  A "step" is one query batch.
`--rows 100000000 --dim 768 --queries 4096` is configs[4] ("C5").
With N > 1 (torchrun, one rank per GPU) the SAME database is row-sharded over the ranks
(strong scaling): every rank searches its shard, one NCCL all-gather of [Q, k] pairs,
merge kernel.  Rank 0 prints one JSON line.

  value  : queries/s with database AND queries resident in HBM (CUDA events, max over ranks)
  e2e    : queries/s through hq.search_stream with the queries in pinned HOST memory and the
           ids/scores read back to the host inside the timed region (host wall clock)
  roofline      : dominant kernel of the step (the rerank contraction) against the burst bf16 peak
  map_index     : the "Hilbert map+index GB/s" half of the metric (fused kernel, HBM roofline)
  cpu_baseline  : the reference's own classes (baseline/_ref) on all host cores on a bounded sample,
                  with the NumPy oracle port beside it; `--impl reference` prints that arm alone
"""
from __future__ import annotations

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

METRIC = "progressive_search_qps_1Mx1536_top10 (with Hilbert map+index GB/s and p50 latency alongside)"
UNIT = "queries/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=1536)
    ap.add_argument("--queries", type=int, default=1024)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-rows", type=int, default=40000)
    ap.add_argument("--cpu-sample-queries", type=int, default=64)
    ap.add_argument("--ref-sample-rows", type=int, default=4000)
    ap.add_argument("--ref-sample-queries", type=int, default=0, help="0 = one query per host core and step")
    ap.add_argument("--bf16-only", action="store_true",
                    help="database without the fp32 rows (bf16 unit rows + index data only): scores carry the bf16 tolerance")
    ap.add_argument("--skip-map-index", action="store_true", help="skip the map+index GB/s pass (large-shard runs)")
    ap.add_argument("--no-graph", action="store_true", help="N > 1, small shards: search eagerly instead of replaying hq.SearchGraph")
    ap.add_argument("--skip-latency", action="store_true", help="skip the single-query latency section")
    return ap.parse_args()


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of this benchmark
    (profiles/r01_traffic.json, written by tools/ncu_summary.py --traffic); None when absent."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            v = json.load(open(os.path.join(ROOT, "profiles", name))).get(kernel)
        except Exception:
            v = None
        if v is not None:
            return v
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "bf16_tflops_burst": p["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_burst": None, "source": "fallback"}


_CLOCK_POLLER = r"""
import sys, time
import pynvml as n
n.nvmlInit()
h = n.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
mx = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
bits = (n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
        n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap)
print("ready", flush=True)
while True:
    sm = float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM))
    r = int(n.nvmlDeviceGetCurrentClocksEventReasons(h))
    print(",".join([str(sm), str(mx), "0"] + ["Active" if r & b else "Not Active" for b in bits]), flush=True)
    time.sleep(0.002)
"""


class ClockSampler:
    """SM clock / throttle reasons during the timed region.  A child process polls NVML every few milliseconds (a query takes
    3-10 ms) from before the region until after it: a timed region of a few tens of milliseconds -- 40 steps of a 125 K-row
    shard -- fell between two samples of `nvidia-smi -lms 100`, and a polling THREAD of this process did not get a query
    through either while the main thread held the interpreter in its launch loop.  `nvidia-smi -lms 100` is the fallback when
    pynvml cannot be used.  Nothing is called from the timed thread: a synchronous NVML query before the end of a short region
    would delay the host by more than the work still queued."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.ready = threading.Event()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis:                                        # torch's device index is a position in CUDA_VISIBLE_DEVICES
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _CLOCK_POLLER, str(self._physical_index())], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            deadline = time.time() + 5.0
            while not self.ready.is_set() and self.proc.poll() is None and time.time() < deadline:
                self.ready.wait(0.05)
            if self.ready.is_set():
                return self
            self.proc.kill()
        except Exception:
            pass
        self.rows = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        proc = self.proc
        for line in proc.stdout:
            if line.startswith("ready"):
                self.ready.set()
                continue
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.02 if self.ready.is_set() else 0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if str(v).lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------
# synthetic data: randn rows, L2-normalised; half the queries are perturbed rows
# --------------------------------------------------------------------------------------
def make_shard(torch, rows, dim, seed, device):
    """fp32 [rows, dim] randn rows, L2-normalised, generated on the device (bench_extra.py and the tools use it)"""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((rows, dim), dtype=torch.float32, device=device)
    step = 1 << 20
    for s in range(0, rows, step):
        e = min(rows, s + step)
        x = torch.randn((e - s, dim), generator=g, device=device)
        x /= x.norm(dim=1, keepdim=True)
        out[s:e] = x
    return out


def make_queries_host(np_db_rows, Q, dim, seed=4321):
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((Q, dim)).astype(np.float32)
    half = min(Q // 2, len(np_db_rows))
    q[:half] = np_db_rows[:half] + 0.1 * rng.standard_normal((half, dim)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q


# --------------------------------------------------------------------------------------
# CPU baseline: the oracle port (NumPy, all BLAS threads) on a bounded sample, extrapolated
# linearly in rows (the algorithm is O(N) per query).
# --------------------------------------------------------------------------------------
def cpu_search_sample(rows, dim, n_queries, k, total_rows):
    from oracle import hilbert_oracle as O
    rng = np.random.default_rng(1234)
    db = rng.standard_normal((rows, dim)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    qs = make_queries_host(db, n_queries, dim)
    n = O.rag_optimal_dimensions(dim)[0]
    t0 = time.perf_counter()
    grids = O.map_to_2d_batch(db, n)
    compact = O.index_c_batch_compact(grids)
    t_index = time.perf_counter() - t0
    levels = O.c_granularity_levels(n)
    rows_l, o = [], 0
    for g in levels:
        r = np.zeros((rows, n), dtype=np.float32)
        r[:, : g * g] = compact[:, o:o + g * g]
        rows_l.append(r)
        o += g * g
    t0 = time.perf_counter()
    for q in qs:
        O.progressive_search(q, db, n, k, db_rows=rows_l)
    t_search = time.perf_counter() - t0
    per_query_full = (t_search / n_queries) * (total_rows / rows)
    bytes_per_row = 4 * dim + 4 * n * n + 4 * compact.shape[1]
    return {"qps": 1.0 / per_query_full, "index_gbs": rows * bytes_per_row / t_index / 1e9,
            "t_search_s": t_search, "t_index_s": t_index}


def cpu_threads():
    try:
        import threadpoolctl
        info = threadpoolctl.threadpool_info()
        return max([i.get("num_threads", 1) for i in info] or [1])
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------
# The reference's OWN classes (baseline/_ref, the unmodified install) on the host cores: per query
# RAGSearchEngineImpl.progressive_hierarchical_search over the enhanced frames -> for every survivor
# _calculate_embedding_cosine_similarity -> stable sort -> top-k, with index rows fed explicitly (the
# SURVEY 8c composition, the one the oracle is pinned against: the reference's height heuristic finds no
# index rows on 1536-D frames).  The reference is single-threaded Python; it gets "all the host threads
# it can use" as one forked worker process per core, each searching whole queries against the shared frames.
# --------------------------------------------------------------------------------------
_REF_STATE = {}


def _ref_one_query(j):
    st = _REF_STATE
    from oracle.reference_loader import rag_filter_with_explicit_rows
    t0 = time.perf_counter()
    surv, eng = rag_filter_with_explicit_rows(st["ref"], st["q_frames"][j], st["frames"], st["n"])
    q_orig = st["q_frames"][j][: st["n"]]
    scored = [(i, eng._calculate_embedding_cosine_similarity(q_orig, st["frames"][i][: st["n"]])) for i in sorted(surv)]
    scored.sort(key=lambda x: x[1], reverse=True)
    return [i for i, _ in scored[: st["k"]]], time.perf_counter() - t0


class ReferenceSample:
    """The unmodified reference classes on a `rows`-row slice of the workload.  Built once (frames through the oracle
    port: the database build is outside the QPS metric), `search()` times one pass of the sample's queries."""

    def __init__(self, rows, dim, n_queries, k, workers=None):
        import multiprocessing as mp
        import warnings
        from oracle import hilbert_oracle as O
        from oracle.reference_loader import load_reference
        warnings.filterwarnings("ignore")
        self.ref = load_reference()
        self.rows, self.k = rows, k
        if self.ref is None:
            return
        self.workers = max(1, workers or (os.cpu_count() or 1))
        self.n_queries = n_queries or self.workers              # default: one query per worker and step
        rng = np.random.default_rng(1234)
        db = rng.standard_normal((rows, dim)).astype(np.float32)
        db /= np.linalg.norm(db, axis=1, keepdims=True)
        qs = make_queries_host(db, self.n_queries, dim)
        n = O.rag_optimal_dimensions(dim)[0]

        def enhanced(x):                                   # enhanced frames [n + L, n]
            grids = O.map_to_2d_batch(x, n)
            compact = O.index_c_batch_compact(grids)
            levels = O.c_granularity_levels(n)
            fr = np.zeros((x.shape[0], n + len(levels), n), dtype=np.float32)
            fr[:, :n] = grids
            o = 0
            for l, g in enumerate(levels):
                w = min(g * g, n)
                fr[:, n + l, :w] = compact[:, o:o + w]
                o += g * g
            return list(fr)
        _REF_STATE.update(ref=self.ref, frames=enhanced(db), q_frames=enhanced(qs), n=n, k=k)
        want, _ = O.progressive_search(qs[0], db, n, k)
        self.want0 = [int(i) for i in want]
        self.pool = mp.get_context("fork").Pool(min(self.workers, self.n_queries)) if self.workers > 1 else None

    def search(self, total_rows):
        t0 = time.perf_counter()
        if self.pool is None:
            res = [_ref_one_query(j) for j in range(self.n_queries)]
        else:
            res = self.pool.map(_ref_one_query, range(self.n_queries), chunksize=1)
        wall = time.perf_counter() - t0
        scale = total_rows / self.rows
        per_query_cpu = float(np.mean([t for _, t in res]))
        return {"qps": self.n_queries / wall / scale, "qps_one_core": 1.0 / (per_query_cpu * scale),
                "workers": min(self.workers, self.n_queries), "wall_s": wall, "agrees_with_port": list(res[0][0]) == self.want0}

    def close(self):
        if getattr(self, "pool", None) is not None:
            self.pool.close()
            self.pool.join()


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on this box's host cores.  When the
    unmodified reference is installed (baseline/_ref) its own classes are timed (kind "reference"); otherwise the
    NumPy oracle port (kind "port").  Each step is a bounded sample, extrapolated linearly in rows."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    times, res, kind = [], None, "reference"
    port = cpu_search_sample(args.cpu_sample_rows, args.dim, args.cpu_sample_queries, args.k, args.rows)
    sampler = ReferenceSample(args.ref_sample_rows, args.dim, args.ref_sample_queries, args.k)
    if sampler.ref is None:
        kind = "port"
    vals = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        if kind == "reference":
            res = sampler.search(args.rows)
        else:
            res = cpu_search_sample(args.cpu_sample_rows, args.dim, args.cpu_sample_queries, args.k, args.rows)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
            vals.append(res["qps"])
    sampler.close()
    value = float(np.mean(vals))
    if kind == "reference":
        sample = (f"unmodified reference classes (baseline/_ref): {sampler.n_queries} queries x {args.ref_sample_rows} rows per step "
                  f"on {res['workers']} forked worker processes (the reference is single-threaded; one query per worker at a time), "
                  f"extrapolated linearly to {args.rows} rows; one core alone: {res['qps_one_core']:.4f} queries/s; "
                  f"top-k ids equal the NumPy port's: {res['agrees_with_port']}")
        cores = res["workers"]
    else:
        sample = (f"NumPy port: {args.cpu_sample_queries} queries x {args.cpu_sample_rows} rows per step, extrapolated linearly to "
                  f"{args.rows} rows")
        cores = cpu_threads()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                             "numpy_port": {"value": port["qps"], "cores": cpu_threads(),
                                            "sample": f"{args.cpu_sample_queries} queries x {args.cpu_sample_rows} rows, extrapolated"}},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "map_index_gbs": port["index_gbs"]}
    print(json.dumps(line))


def cpu_baseline_subprocess(args):
    """The CPU leg of OUR arm (rank 0, N = 1): one step of `--impl reference` in a fresh process (its worker pool forks;
    forking this process after CUDA initialisation would not be safe).  Returns its `cpu_baseline` object."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--rows", str(args.rows), "--dim", str(args.dim), "--queries", str(args.queries), "--k", str(args.k),
           "--cpu-sample-rows", str(args.cpu_sample_rows), "--cpu-sample-queries", str(args.cpu_sample_queries),
           "--ref-sample-rows", str(args.ref_sample_rows), "--ref-sample-queries", str(args.ref_sample_queries)]
    try:
        t0 = time.perf_counter()
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=600, cwd=ROOT,
                             env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        line = json.loads(out.stdout.strip().splitlines()[-1])
        cb = line["cpu_baseline"]
        cb["map_index_gbs"] = line.get("map_index_gbs")
        cb["wall_s"] = round(time.perf_counter() - t0, 1)
        return cb
    except Exception as ex:                                   # reported, never hidden
        return {"error": f"{type(ex).__name__}: {ex}"[:300]}


def workload_name(args):
    tag = {(1_000_000, 1536): "C2", (100_000_000, 768): "C5", (25_000, 1536): "C1"}.get((args.rows, args.dim), "custom")
    return f"{tag}: {args.rows}x{args.dim} fp32 embeddings, {args.queries}-query batches, progressive top-{args.k}"


def workload_config(args, world):
    """The `config` object of BOTH arms (ours and `--impl reference`): same keys, same values."""
    per = -(-args.rows // world)
    return {"workload": workload_name(args), "rows": args.rows, "dim": args.dim, "queries_per_step": args.queries, "k": args.k,
            "filter_scope": "shard", "rows_per_gpu": per,
            "sharding": f"row-sharded x{world}, one NCCL all-gather of [Q,k]" if world > 1 else "single shard",
            "l2": "database and bit planes exceed the 126 MB L2 many times over: no flush needed between steps"}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import hilbert_quantization_b200 as hq
    from hilbert_quantization_b200 import search as S
    from hilbert_quantization_b200._lib import lib
    from hilbert_quantization_b200.distributed import allgather_merge, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout; the contract is ONE JSON line there, so the
        # process-level stdout is pointed at stderr until the communicator exists.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    pk = peaks()

    # ---- database shard (rows [lo, hi) of the global database), generated on the device ----
    lo, hi = shard_bounds(args.rows, world, rank)
    rows = hi - lo
    n = hq.rag_optimal_dimensions(args.dim)[0]
    gen_chunk = 1 << 20

    def shard_chunks():
        """fp32 row blocks of this rank's shard (seed 1234 + rank, L2-normalised randn), one block at a time"""
        g = torch.Generator(device=device).manual_seed(1234 + rank)
        for s in range(0, rows, gen_chunk):
            e = min(rows, s + gen_chunk)
            x = torch.randn((e - s, args.dim), generator=g, device=device)
            x /= x.norm(dim=1, keepdim=True)
            yield x

    emb = None
    if args.bf16_only:
        first = next(shard_chunks())
        head_rows, mi_src = first[: args.queries // 2], first
    else:
        emb = torch.empty((rows, args.dim), dtype=torch.float32, device=device)
        at = 0
        for x in shard_chunks():
            emb[at:at + x.shape[0]] = x
            at += x.shape[0]
        del x
        head_rows, mi_src = emb[: args.queries // 2], emb[: 1 << 20]
    head = head_rows.cpu().numpy() if rank == 0 else None
    q_host = make_queries_host(head if head is not None else np.zeros((0, args.dim), np.float32), args.queries, args.dim)
    q_pinned = torch.from_numpy(q_host).pin_memory()
    if world > 1:                                    # every rank must search the same queries
        qd = q_pinned.to(device)
        dist.broadcast(qd, 0)
        q_pinned.copy_(qd.cpu())
    q_dev = q_pinned.to(device)

    # ---- "Hilbert map+index GB/s": fused map_to_2d + variant-C index over (the first 1 Mi rows of) this shard ----
    map_index = None
    if not args.skip_map_index:
        L_idx = sum(min(g * g, n) for g in hq.index.plans.c_levels(n))
        bytes_per_row = 4 * args.dim + 4 * n * n + 4 * L_idx        # read D, write grid, write index
        mi_rows = int(mi_src.shape[0])
        chunk = min(mi_rows, 262144)
        grids_buf = torch.empty((chunk, n, n), dtype=torch.float32, device=device)
        from hilbert_quantization_b200.index import fused_pass, plans
        plan, widths, ml = plans.c_plan(n, "compact")
        idx_buf = torch.empty((chunk, len(plan)), dtype=torch.float32, device=device)

        def map_index_pass():
            for s in range(0, mi_rows, chunk):
                e = min(mi_rows, s + chunk)
                fused_pass(mi_src[s:e], 0, n, args.dim, plan=plan, plan_key=("C", n, "compact"), min_level=ml,
                           grid_out=grids_buf[: e - s].view(e - s, -1), idx_out=idx_buf[: e - s])
        for _ in range(max(3, args.warmup)):
            map_index_pass()
        torch.cuda.synchronize()
        mi_ms = []
        for _ in range(max(3, args.steps)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            map_index_pass()
            e1.record()
            torch.cuda.synchronize()
            mi_ms.append(e0.elapsed_time(e1))
        mi_t = torch.tensor([float(np.mean(mi_ms))], device=device)
        if world > 1:
            dist.all_reduce(mi_t, op=dist.ReduceOp.MAX)
        mi_kernel_gbs = mi_rows * bytes_per_row / (float(np.mean(mi_ms)) * 1e-3) / 1e9
        map_index = {"value": world * mi_rows * bytes_per_row / (mi_t.item() * 1e-3) / 1e9, "unit": "GB/s",
                     "bytes_per_embedding": bytes_per_row, "rows_per_gpu": mi_rows,
                     "ms_per_pass": mi_t.item(), "launches_per_pass": (mi_rows + chunk - 1) // chunk,
                     "roofline": {"kernel": "k_item_pass_bulk<0,6> (fused map_to_2d + index pyramid, bulk-copy loads and stores)",
                                  "bound": "hbm", "achieved": mi_kernel_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                  "frac": mi_kernel_gbs / pk["hbm_gbs"],
                                  "traffic": ncu_traffic("k_item_pass_bulk") if chunk == 262144 else None,
                                  "peak_source": pk["source"]}}
        del grids_buf, idx_buf
        # the same rows through the fused map + index + uint8 quantise kernel (north star (2)): uint8 enhanced frames out
        rows_c = len(plans.c_plan(n, "rows")[0]) // n
        q_bytes_per_row = 4 * args.dim + n * n + rows_c * n + 8

        def quant_pass():
            for s in range(0, mi_rows, chunk):
                hq.map_index_quantize(mi_src[s:min(mi_rows, s + chunk)], n, variant="C")
        for _ in range(max(3, args.warmup)):
            quant_pass()
        torch.cuda.synchronize()
        mq_ms = []
        for _ in range(max(3, args.steps)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            quant_pass()
            e1.record()
            torch.cuda.synchronize()
            mq_ms.append(e0.elapsed_time(e1))
        mq = float(np.mean(mq_ms))
        map_index["quantized"] = {"value": mi_rows * q_bytes_per_row / (mq * 1e-3) / 1e9, "unit": "GB/s", "ms_per_pass": mq,
                                  "rows_per_s": mi_rows / (mq * 1e-3), "bytes_per_embedding": q_bytes_per_row,
                                  "kernel": "k_item_pass_bulk<0,6,quant> (map_to_2d + variant-C index + embed + uint8 min/max normalise, "
                                            "one launch; replaces the map+index pass AND the separate quantise pass over the frames)",
                                  "roofline": {"bound": "hbm", "achieved": mi_rows * q_bytes_per_row / (mq * 1e-3) / 1e9,
                                               "peak": pk["hbm_gbs"], "unit": "GB/s",
                                               "frac": mi_rows * q_bytes_per_row / (mq * 1e-3) / 1e9 / pk["hbm_gbs"],
                                               "note": "issue bound, not HBM bound: the pass moves 2.2x fewer bytes than the grid-writing "
                                                       "pass in a similar time (see DESIGN.md section 5)"}}
    del mi_src

    # ---- database build (untimed) ----
    if args.bf16_only:
        db = hq.EmbeddingDatabase.from_chunks(shard_chunks(), rows, args.dim, n=n, device=device, id_base=lo)
    else:
        db = hq.EmbeddingDatabase(emb, n=n, device=device, id_base=lo)
    torch.cuda.synchronize()

    # N > 1: the all-gather + merge of batch i runs on a communication stream under the search of batch i + 1
    # (hq.distributed.MergePipeline); `wait` makes the main stream wait for it (latency measurements, the last batch)
    from hilbert_quantization_b200.distributed import MergePipeline
    pipe = MergePipeline(device) if world > 1 else None

    # Small shards (a GPU's share of C2 over 4 or 8 GPUs) are launch bound: ~45 launches for 0.8 ms of kernels.  There the
    # throughput loop replays the search as a CUDA graph (hq.SearchGraph, the public latency / small-batch path); two
    # graphs alternate so that the merge of batch i can still read its output buffers while batch i + 1 is searched.
    graphs = None
    if world > 1 and rows <= 500_000 and not args.no_graph:
        try:
            graphs = [hq.SearchGraph(db, args.queries, args.k) for _ in range(2)]
        except Exception as ex:                       # reported in the line, the eager path runs instead
            graphs = None
            print(f"[bench] SearchGraph unavailable, eager search: {type(ex).__name__}: {ex}"[:300], file=sys.stderr)
    graph_turn = [0]

    def step(queries, wait=True):
        if graphs is not None and tuple(queries.shape) == (args.queries, args.dim):
            pipe.reserve()
            g = graphs[graph_turn[0] & 1]
            graph_turn[0] += 1
            ids, sc = g.search(queries, sync=False)
        else:
            ids, sc = hq.search_batch(db, queries, args.k)
        if world > 1:
            ids, sc, _ = pipe.submit(ids, sc, args.k)
            if wait:
                pipe.drain()
        return ids, sc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        step(q_dev)
    barrier()
    lib.hq_launch_count(1)
    lib.hq_kernel_timing(1)                           # CUDA events around the launches of the four search kernels (roofline)
    S.PHASE_TIMER = S.PhaseTimer()
    step_ms = []
    with ClockSampler(local) as clocks:
        t_all0 = torch.cuda.Event(enable_timing=True)
        t_all1 = torch.cuda.Event(enable_timing=True)
        t_all0.record()
        for _ in range(args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(q_dev, wait=False)
            e1.record()
            step_ms.append((e0, e1))
        if pipe is not None:
            pipe.drain()                              # the timed region ends when the last batch's merged result exists
        t_all1.record()
        barrier()
    total_ms = t_all0.elapsed_time(t_all1)
    import ctypes as _C
    k_ms, k_n = (_C.c_float * 4)(), (_C.c_int32 * 4)()
    lib.hq_kernel_timing_read(k_ms, k_n, 4)
    kernel_ms = {name: (float(k_ms[i]) / args.steps, int(k_n[i]) // max(1, args.steps))
                 for i, name in enumerate(("k_rerank_tc", "k_filter_bits_tc", "k_filter_cascade", "k_rerank_tc_merge"))}
    launches = int(lib.hq_launch_count(1))
    phases = S.PHASE_TIMER.totals_ms()
    S.PHASE_TIMER = None
    kernels_from = "timed region"
    if graphs is not None:
        # a graph replay does not pass through the launchers: the per-kernel and per-phase times (and the launch count of
        # one step) come from a few eager steps AFTER the timed region; the replayed step launches the same kernels
        for g in graphs:
            if g.dense is not None and int(g.dense.item()) == 0:
                raise RuntimeError("a query of the batch has a sparse index row: the graph path does not apply")
        n_e = max(1, min(5, args.steps))
        lib.hq_launch_count(1)
        lib.hq_kernel_timing(1)
        S.PHASE_TIMER = S.PhaseTimer()
        for _ in range(n_e):
            ids_e, sc_e = hq.search_batch(db, q_dev, args.k)
            pipe.submit(ids_e, sc_e, args.k)
        pipe.drain()
        torch.cuda.synchronize()
        lib.hq_kernel_timing_read(k_ms, k_n, 4)
        kernel_ms = {name: (float(k_ms[i]) / n_e, int(k_n[i]) // n_e)
                     for i, name in enumerate(("k_rerank_tc", "k_filter_bits_tc", "k_filter_cascade", "k_rerank_tc_merge"))}
        launches = int(lib.hq_launch_count(1)) * args.steps // n_e
        phases = {k_: v * args.steps / n_e for k_, v in S.PHASE_TIMER.totals_ms().items()}
        S.PHASE_TIMER = None
        kernels_from = f"{n_e} eager steps after the timed region (the timed steps replay the same kernels as a CUDA graph)"
        barrier()
    per_step = [a.elapsed_time(b) for a, b in step_ms]
    tt = torch.tensor([total_ms], device=device)
    per_rank_ms = [total_ms / args.steps]
    if world > 1:
        allt = torch.empty(world, device=device)
        dist.all_gather_into_tensor(allt, tt.clone())
        per_rank_ms = [float(x) / args.steps for x in allt.cpu()]
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = tt.item()
    ms_per_step = total_ms / args.steps
    qps = args.queries * args.steps / (total_ms * 1e-3)

    # ---- end to end: queries from pinned host memory, results back on the host ----
    out_ids = torch.empty((args.queries, args.k), dtype=torch.int64).pin_memory()
    out_sc = torch.empty((args.queries, args.k), dtype=torch.float32).pin_memory()

    # hq.search_stream: the public call for host-resident query batches.  Every step copies its 1024 queries from pinned
    # host memory and reads its ids / scores back on the host; the copy of step i + 1 runs on a copy stream under the search
    # of step i (two slots), the host consumes step i's results while step i + 1 runs.
    post = (lambda i_, s_: allgather_merge(i_, s_, args.k)) if world > 1 else None
    post_stream = pipe.stream if pipe is not None else None

    e2e_stamps = []

    def e2e_run(n_steps):
        acc = 0.0
        e2e_stamps.clear()
        for ids_h, sc_h in hq.search_stream(db, (q_pinned for _ in range(n_steps)), args.k, post=post, post_stream=post_stream):
            out_ids.copy_(ids_h)
            out_sc.copy_(sc_h)
            acc += float(out_sc[0, 0])
            e2e_stamps.append(time.perf_counter())
        return acc
    e2e_run(args.warmup)
    barrier()
    import gc
    gc_ms = [0.0, 0]
    gc_t = [0.0]

    def gc_probe(phase, info):                        # diagnostic: Python garbage collections inside the timed region
        if phase == "start":
            gc_t[0] = time.perf_counter()
        else:
            gc_ms[0] += (time.perf_counter() - gc_t[0]) * 1e3
            gc_ms[1] += 1
    gc.callbacks.append(gc_probe)
    t0 = time.perf_counter()                          # host wall clock: the e2e figure includes everything the caller waits for
    e2e_run(args.steps)                               # (returns after the last batch's results are on the host)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    gc.callbacks.remove(gc_probe)
    gaps = np.diff([t0] + e2e_stamps) * 1e3              # host time between consecutive result batches
    barrier()
    te = torch.tensor([e2e_ms], device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_qps = args.queries * args.steps / (te.item() * 1e-3)

    latency = None
    if not args.skip_latency:
        # ---- single-query latency on the same database (the "p50 latency at 1M x 1536" part of the metric) ----
        q1 = q_dev[:1].contiguous()
        for _ in range(5):
            step(q1)
        barrier()
        q1_ms = []
        for _ in range(50):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(q1)
            e1.record()
            e1.synchronize()
            q1_ms.append(e0.elapsed_time(e1))
        # the same query through SearchGraph (the search captured once as a CUDA graph: one launch instead of fifteen)
        g1_ms, g1_err = [], None
        try:
            graph1 = hq.SearchGraph(db, 1, args.k)

            def gstep(queries):
                ids, sc = graph1.search(queries)
                if world > 1:
                    ids, sc = allgather_merge(ids, sc, args.k)
                return ids, sc
            for _ in range(5):
                gstep(q1)
            barrier()
            ids_g, sc_g = [t.clone() for t in gstep(q1)]
            ids_e, sc_e = step(q1)
            if not (torch.equal(ids_g, ids_e) and torch.equal(sc_g, sc_e)):
                raise RuntimeError("SearchGraph and search_batch disagree")
            for _ in range(50):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gstep(q1)
                e1.record()
                e1.synchronize()
                g1_ms.append(e0.elapsed_time(e1))
        except Exception as ex:                                        # reported in the line, never hidden
            g1_err = f"{type(ex).__name__}: {ex}"[:200]
            g1_ms = [float("nan")]
        q1_t = torch.tensor([float(np.median(q1_ms)), float(np.percentile(q1_ms, 99)),
                             float(np.median(g1_ms)), float(np.percentile(g1_ms, 99))], device=device)
        if world > 1:
            dist.all_reduce(q1_t, op=dist.ReduceOp.MAX)
        q1_p50, q1_p99, g1_p50, g1_p99 = [float(x) for x in q1_t.cpu()]

        latency = {"p50": q1_p50, "p99": q1_p99, "note": "one query per call, same database, device timed",
                   "cuda_graph": ({"p50": g1_p50, "p99": g1_p99, "note": "hq.SearchGraph(db, 1, k).search(q): the same "
                                   "kernels replayed as one CUDA graph, results checked equal"}
                                  if g1_err is None else {"error": g1_err})}

    # ---- sanity: perturbed queries should surface their source row when it survives the filter ----
    ids, sc = step(q_dev)
    torch.cuda.synchronize()

    if rank == 0:
        phase_ms = phases.get("rerank_gemm", 0.0) / args.steps    # GEMM + shortlist merge + guard kernels
        gemm_ms = kernel_ms["k_rerank_tc"][0] or phase_ms          # the GEMM kernel alone (events around its launches)
        flops = 2.0 * args.queries * rows * args.dim               # per step, this rank's shard (one launch per query chunk)
        achieved_tf = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        # roofline denominator: the BURST cuBLAS bf16 figure (the timed region is a fraction of a second; the sustained
        # figure, which this kernel exceeded in round 1, is kept as a side field)
        peak_tf = pk.get("bf16_tflops_burst") or pk["bf16_tflops"]
        cfg = workload_config(args, world)
        line = {
            "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16 tcgen05 contraction (fp32 accumulate) + exact fp32 re-score of the shortlist; filter: tf32 hi/lo split",
            "data": "synthetic",
            "config": cfg,
            "details": {"grid": f"{n}x{n} Hilbert grids + variant-C index", "rows_this_gpu": rows,
                        "database": "bf16 unit rows only (no fp32 copy)" if args.bf16_only else "fp32 rows + bf16 unit rows",
                        "rerank": "tcgen05 bf16 contraction + fused mask/top-16 epilogue, exact fp32 re-score of the shortlist "
                                  "with a shortlist-sufficiency guard (flagged queries re-scored exactly over all survivors)",
                        "filter": "window mode: tcgen05 tf32 (hi/lo split) sample pass -> predicted cut windows -> window pass (alive "
                                  "plane + counters + lists of the window rows) -> exact ranking of the window rows; mispredicted "
                                  "queries are redone by the fallback pass (full-threshold planes + generic gather cascade)"},
            "p50_ms": float(np.median(per_step)), "p99_ms": float(np.percentile(per_step, 99)),
            "single_query_latency_ms": latency,
            "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": int(q_pinned.numel() * 4),
                    "d2h_bytes_per_step": int(out_ids.numel() * 8 + out_sc.numel() * 4),
                    "clock": "time.perf_counter around the loop (barrier + synchronize on both sides), max over ranks",
                    "result_gaps_ms": {"first": float(gaps[0]), "median": float(np.median(gaps[1:])) if len(gaps) > 1 else None,
                                       "max": float(gaps[1:].max()) if len(gaps) > 1 else None,
                                       "argmax": int(1 + np.argmax(gaps[1:])) if len(gaps) > 1 else None,
                                       "gc_ms": gc_ms[0], "gc_collections": gc_ms[1]}},
            "gpu_launches": launches,
            "per_rank_ms_per_step": per_rank_ms,
            "phases_ms_per_step": {k: v / args.steps for k, v in phases.items()},
            "kernels_ms_per_step": {k: {"ms": v[0], "launches": v[1]} for k, v in kernel_ms.items()},
            "kernels_measured_in": kernels_from,
            "search_replayed_as_cuda_graph": graphs is not None,
            "roofline": {"kernel": "k_rerank_tc<16> (Q x N x D cosine contraction, bf16 tcgen05, fused mask / top-16 epilogue)", "bound": "tensor",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": (achieved_tf / peak_tf) if achieved_tf else None,
                         "traffic": ncu_traffic("k_rerank_tc") if world == 1 and args.rows == 1_000_000 else None,
                         "peak_source": pk["source"] + " (burst cuBLAS bf16)",
                         "frac_of_sustained_peak": (achieved_tf / pk["bf16_tflops"]) if achieved_tf else None,
                         "algorithmic_flops_per_launch": flops / max(1, kernel_ms["k_rerank_tc"][1]),
                         "launches_per_step": kernel_ms["k_rerank_tc"][1], "kernel_ms_per_step": gemm_ms,
                         "timing": "CUDA events on the launching stream around every k_rerank_tc launch of the timed region",
                         "frac_with_merge_and_guard": (flops / (phase_ms * 1e-3) / 1e12 / peak_tf) if phase_ms > 0 else None,
                         "note": "largest single kernel of the step; the coarse filter (tcgen05 tf32 window pass + window "
                                 "cascade) is issue bound, see DESIGN.md section 5 (K5w) and profiles/"},
            "clocks": clocks.summary(),
            "top1_hit_rate_perturbed": float((ids[: args.queries // 2, 0].cpu().numpy() == np.arange(args.queries // 2)).mean())
            if lo == 0 else None,
        }
        if map_index is not None:
            line["map_index"] = map_index
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_subprocess(args)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
