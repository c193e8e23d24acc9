#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs 0, 2, 3 and a 768-D shard of config 4) on one
B200.  Prints one JSON line per workload; each carries its roofline against the measured
peaks and a size-independent parity property checked at full size.

    python bench_extra.py [--only c1,c3,c4,c5]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import hilbert_quantization_b200 as hq                      # noqa: E402
from hilbert_quantization_b200.index import fused_pass, plans   # noqa: E402
from bench import ClockSampler, make_queries_host, make_shard, peaks       # noqa: E402


def timed(fn, warmup=3, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts)), float(np.min(ts))


def c1(pk):
    """README benchmark: 25K x 1536 -> 64x64 + indices, ONE query, top-10 (latency)."""
    dev = torch.device("cuda")
    emb = make_shard(torch, 25_000, 1536, 1234, dev)
    db = hq.EmbeddingDatabase(emb)
    q = torch.from_numpy(make_queries_host(emb[:1].cpu().numpy(), 2, 1536)[:1]).to(dev)
    ms, best = timed(lambda: hq.search_batch(db, q, 10), warmup=5, iters=50)
    qp = torch.from_numpy(q.cpu().numpy()).pin_memory()
    out_i = torch.empty((1, 10), dtype=torch.int64).pin_memory()

    def e2e():
        ids, sc = hq.search_batch(db, qp.to(dev, non_blocking=True), 10)
        out_i.copy_(ids, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        e2e()
    wall = (time.perf_counter() - t0) / 50 * 1e3
    # the same query through the search captured as a CUDA graph, from pinned host memory to ids on the host
    sg = hq.SearchGraph(db, 1, 10)
    ids_e, sc_e = hq.search_batch(db, q, 10)
    ids_g, sc_g = sg.search(q)
    same = bool(torch.equal(ids_g, ids_e) and torch.equal(sc_g, sc_e))
    g_ms, g_best = timed(lambda: sg.search(q), warmup=5, iters=50)

    def e2e_graph():
        ids, sc = sg.search(qp)
        out_i.copy_(ids, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(5):
        e2e_graph()
    t0 = time.perf_counter()
    for _ in range(50):
        e2e_graph()
    g_wall = (time.perf_counter() - t0) / 50 * 1e3
    return {"workload": "C1: 25K x 1536, 1 query, progressive top-10", "latency_ms_device": ms, "latency_ms_best": best,
            "latency_ms_e2e_host": wall, "cuda_graph": {"latency_ms_device": g_ms, "latency_ms_best": g_best,
                                                        "latency_ms_e2e_host": g_wall, "same_results": same}, "reference_published_ms": 4.6, "unit": "ms"}


def c3(pk, rows=10_000_000, chunk=2_000_000):
    """map_from_2d round trip of 10M x 1024 (32x32), bitwise check at full size."""
    dev = torch.device("cuda")
    mapper = hq.HilbertCurveMapper()
    x = make_shard(torch, chunk, 1024, 7, dev)
    grids = mapper.map_to_2d_batch(x, 32)
    out = torch.empty_like(x)
    n_chunks = rows // chunk
    ms, best = timed(lambda: mapper.map_from_2d_batch(grids, out=out), warmup=3, iters=n_chunks)
    ok = bool(torch.equal(out, x))
    # XOR-fold checksum of the whole round trip (permutation invariant)
    fold_in = int(x.view(torch.int32).sum(dtype=torch.int64).item())
    fold_out = int(out.view(torch.int32).sum(dtype=torch.int64).item())
    bytes_per = 2 * 1024 * 4
    gbs = chunk * bytes_per / (ms * 1e-3) / 1e9
    return {"workload": f"C3: map_from_2d of {rows} x 1024 (32x32), timed as {n_chunks} launches of {chunk} grids",
            "value": gbs, "unit": "GB/s", "ms_total": ms * n_chunks, "bit_exact_round_trip": ok and fold_in == fold_out,
            "roofline": {"kernel": "k_item_pass<1,0,5> (unmap)", "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"],
                         "frac": gbs / pk["hbm_gbs"], "unit": "GB/s", "bytes_per_embedding": bytes_per}}


def c4(pk):
    """Qwen2.5-0.5B sized fp32 parameter stream (494,032,768 values) -> 30 grids of 4096^2 + variant-C indices."""
    dev = torch.device("cuda")
    total, n = 494_032_768, 4096
    cells = n * n
    grids_n = -(-total // cells)
    g = torch.Generator(device=dev).manual_seed(0)
    stream = torch.empty(grids_n * cells, dtype=torch.float32, device=dev)
    stream[:total].normal_(0.0, 0.02, generator=g)
    stream[total:] = 0
    plan, widths, ml = plans.c_plan(n, "compact")
    out_grids = torch.empty((grids_n, n, n), dtype=torch.float32, device=dev)
    idx = torch.empty((grids_n, len(plan)), dtype=torch.float32, device=dev)
    tail = total - (total // cells) * cells
    params = stream[:total]                                   # the model's parameters, NOT padded

    def run():
        # all 30 grids in one launch; the kernel zero-fills the tail of the last grid (fill 0.447)
        hq.map_parameter_stream(params, n, variant="C", grid_out=out_grids, idx_out=idx)
    ms, best = timed(run, warmup=2, iters=5)
    bytes_total = 4 * total + 4 * grids_n * cells + 4 * idx.numel()
    gbs = bytes_total / (ms * 1e-3) / 1e9
    # parity at full size: inverse map returns the stream bit-exactly; indices against run means of the stream
    back = hq.HilbertCurveMapper().map_from_2d_batch(out_grids)
    ok = bool(torch.equal(back.view(-1)[:total], stream[:total])) and bool((back.view(-1)[total:] == 0).all())
    lvl0 = stream.view(grids_n, cells // 4096, 4096).double().mean(dim=2)           # g = 64 row: 4096-run means
    err = float((idx[:, :4096].double() - lvl0).abs().max())
    return {"workload": f"C4: {total} fp32 parameters -> {grids_n} grids of 4096x4096 (last at fill {tail / cells:.3f}) + variant-C indices",
            "value": gbs, "unit": "GB/s", "ms_total": ms, "bit_exact_inverse": ok, "index_max_abs_err_vs_fp64": err,
            "roofline": {"kernel": "k_tile_pass_bulk<0> + k_pyramid_top<0> (hq_map_index_stream: one launch over 29 full grids + 1 partial grid)", "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"],
                         "frac": gbs / pk["hbm_gbs"], "unit": "GB/s", "algorithmic_bytes": bytes_total}}


def c5(pk, rows=12_500_000, queries=4096):
    """One shard of config 4 (100M x 768 over 8 GPUs = 12.5M rows per GPU), 4096-query batches."""
    dev = torch.device("cuda")
    emb = make_shard(torch, rows, 768, 1234, dev)
    t0 = time.perf_counter()
    db = hq.EmbeddingDatabase(emb)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    q = torch.from_numpy(make_queries_host(emb[: queries // 2].cpu().numpy(), queries, 768)).to(dev)
    ms, best = timed(lambda: hq.search_batch(db, q, 10), warmup=2, iters=3)
    flop = 2.0 * queries * rows * 768
    return {"workload": f"C5 shard: {rows} x 768 (32x32), {queries}-query batches, progressive top-10 on one GPU",
            "value": queries / (ms * 1e-3), "unit": "queries/s", "ms_per_batch": ms, "db_build_s": build_s,
            "rerank_gemm_tflops_if_whole_step": flop / (ms * 1e-3) / 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c1,c3,c4,c5")
    args = ap.parse_args()
    pk = peaks()
    for name in args.only.split(","):
        fn = {"c1": c1, "c3": c3, "c4": c4, "c5": c5}[name]
        try:
            with ClockSampler(0) as clocks:                      # nvidia-smi clocks / throttle reasons while the workload runs
                res = fn(pk)
            res["clocks"] = clocks.summary()
        except Exception as e:                                   # keep the other workloads running
            res = {"workload": name, "error": repr(e)[:400]}
        res["name"] = name
        print(json.dumps(res), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
