"""CPU suite: host logic, planning, the C-ABI surface (load + exported symbols, no compute),
and the multi-rank merge path over gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import hilbert_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from hilbert_quantization_b200 import _lib
    header = open(os.path.join(ROOT, "include", "hq_b200.h")).read()
    declared = set(re.findall(r"\b(hq_[a-z0-9_]+)\s*\(", header))
    declared -= {"hq_index_layout"}
    assert len(declared) >= 24
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/hq_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert _lib.lib.hq_version() == 1
    assert isinstance(_lib.last_error(), str)


def test_ctypes_signatures_match_the_header():
    """Every prototype of include/hq_b200.h against the argtypes list the Python side binds it with: parameter count and
    parameter kind (a wrong count or width would shift or truncate every later argument of a call silently)."""
    import ctypes
    from hilbert_quantization_b200 import _lib
    header = open(os.path.join(ROOT, "include", "hq_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)                       # comments quote prototypes too
    protos = re.findall(r"\b(?:int|int64_t|const char\s*\*)\s+(hq_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    assert len(protos) >= 60
    seen = set()
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert name in _lib.SIGNATURES, name
        sig = _lib.SIGNATURES[name][1]
        assert len(sig) == n, f"{name}: header has {n} parameters, ctypes binds {len(sig)}"
        for i, (par, t) in enumerate(zip([x.strip() for x in params.split(",")] if n else [], sig)):
            if "*" in par:
                ok = t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "_type_")          # POINTER(IndexLayout)
            elif re.match(r"(const\s+)?int64_t\b", par):
                ok = t is ctypes.c_int64
            elif re.match(r"(const\s+)?int\b", par):
                ok = t in (ctypes.c_int32, ctypes.c_int)
            elif re.match(r"(const\s+)?float\b", par):
                ok = t is ctypes.c_float
            elif re.match(r"(const\s+)?double\b", par):
                ok = t is ctypes.c_double
            else:
                ok = False
            assert ok, f"{name}: parameter {i} `{par}` is bound as {t}"
        seen.add(name)
    assert {"hq_rerank_sparse_topk", "hq_filter_fast_rows", "hq_filter_rows_pack", "hq_rerank_topk_unit_bf16"} <= seen


def test_bench_clock_sampler_degrades_without_nvml():
    """bench.ClockSampler on a box without a GPU: no samples, the summary keeps its keys, no long wait."""
    import time
    sys.path.insert(0, ROOT)
    import bench
    t0 = time.time()
    with bench.ClockSampler(0) as c:
        time.sleep(0.02)
    out = c.summary()
    assert set(out) == {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
    assert time.time() - t0 < 4.0


def test_argument_validation_without_gpu():
    """HQ_EINVAL paths return before any CUDA call, so they are checkable on the CPU box."""
    from hilbert_quantization_b200._lib import lib, last_error, HQ_EINVAL
    assert lib.hq_d2xy_batch(3, 0, 1, None, None, None) == HQ_EINVAL and "power of 2" in last_error()
    assert lib.hq_map_to_2d(None, 1, 20, 20, 4, 4, None, 16, None) == HQ_EINVAL and "Too many parameters" in last_error()
    assert lib.hq_map_to_2d(None, 1, 4, 4, 4, 3, None, 16, None) == HQ_EINVAL
    assert lib.hq_topk_from_scores(None, 0, 0, 1, 0, 0, None, None, None) == HQ_EINVAL


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import hilbert_quantization_b200 as hq
    with pytest.raises(hq.HQLibraryError, match="no CPU fallback"):
        hq.HilbertCurveMapper().map_to_2d(np.arange(4, dtype=np.float32), (2, 2))
    with pytest.raises(hq.HQLibraryError):
        hq.EmbeddingDatabase(np.zeros((4, 16), dtype=np.float32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "hilbert_quantization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def _emulate(img, plan, mode):
    """CPU model of the kernel's pyramid + gather (fp32 pairwise tree / fp64 strict)."""
    stream = O.map_from_2d(img)
    cur = stream.astype(np.float64 if mode else np.float32)
    levels = []
    while len(cur) >= 4:
        g = cur.reshape(-1, 4)
        if mode:
            cur = (((g[:, 0] + g[:, 1]) + g[:, 2]) + g[:, 3]) * 0.25
        else:
            cur = (((g[:, 0] + g[:, 1]) + (g[:, 2] + g[:, 3])) * np.float32(0.25)).astype(np.float32)
        levels.append(cur)
    addr = np.concatenate([img.reshape(-1).astype(np.float64)] + [l.astype(np.float64) for l in levels])
    return np.where(plan >= 0, addr[np.maximum(plan, 0)], 0.0)


@pytest.mark.parametrize("n,D", [(4, 16), (8, 50), (16, 200), (32, 768), (64, 1536), (128, 16384)])
def test_gather_plans_reproduce_the_oracle(n, D):
    from hilbert_quantization_b200 import plans as P
    rng = np.random.default_rng(n)
    img = O.map_to_2d(rng.standard_normal(D).astype(np.float32), (n, n))
    plan, widths, ml = P.c_plan(n, "compact")
    assert widths == [min(g * g, n) for g in O.c_granularity_levels(n)]
    assert np.abs(_emulate(img, plan, 0) - O.index_c_batch_compact(img[None])[0]).max() <= 2e-7
    rows_plan, _, _ = P.c_plan(n, "rows")
    assert np.abs(_emulate(img, rows_plan, 0).reshape(-1, n) - O.index_c(img)[n:]).max() <= 2e-7
    for S in (n, 7, 100, 1024):
        pb, _ = P.b_plan(n, S)
        assert np.array_equal(_emulate(img, pb, 1), O.index_b(img, S))            # bit exact
        pa, _ = P.a_plan(n, S)
        assert np.abs(_emulate(img, pa, 0) - O.index_a(img, S)).max() <= 2e-7
    assert P.a_allocation(64) == O.a_level_allocation(64)
    assert P.core_levels(64, 64) == O.core_parse_levels(64, 64)


def test_host_planning_matches_oracle_scalars():
    from hilbert_quantization_b200 import plans as P
    from hilbert_quantization_b200.dimension import PowerOf4DimensionCalculator, rag_optimal_dimensions
    from hilbert_quantization_b200.search import make_layout, rag_ratio, rag_threshold
    calc = PowerOf4DimensionCalculator()
    for c in (1, 4, 5, 768, 1024, 1025, 1536, 4096, 16385, 16777216):
        assert calc.calculate_optimal_dimensions(c) == O.optimal_dimensions(c)
        assert rag_optimal_dimensions(c) == O.rag_optimal_dimensions(c)
    with pytest.raises(ValueError, match="below minimum"):
        calc.calculate_padding_strategy(1536, (64, 64))
    assert calc.calculate_padding_strategy(768, (32, 32)).efficiency_ratio == 0.75
    with pytest.raises(ValueError, match="must be positive"):
        calc.calculate_optimal_dimensions(0)
    for w in (2, 4, 8, 16, 32, 64, 100, 128, 1024, 4096):
        assert P.c_levels(w) == O.c_granularity_levels(w)
    for l in range(6):
        assert rag_threshold(l) == O.rag_threshold(l) and rag_ratio(l) == O.rag_ratio(l)
    for n in (2, 4, 8, 16, 32, 64):
        d = np.arange(n * n)
        x, y = P._d2xy(n, d)
        ox, oy = O.d2xy(n, d)
        assert np.array_equal(x, ox) and np.array_equal(y, oy) and np.array_equal(P._xy2d(n, x, y), d)
    lay, levels = make_layout(64, 1536)
    assert levels == [8, 4, 2] and lay.Lsum == 84 and list(lay.lvl_keff)[:3] == [24, 6, 4]
    lay, _ = make_layout(32, 768)
    assert lay.Lsum == 20 and list(lay.lvl_keff)[:2] == [12, 4]
    lay, _ = make_layout(32, 200)
    assert list(lay.lvl_keff)[:2] == [4, 1]


def test_shard_bounds_cover_rows():
    from hilbert_quantization_b200.distributed import shard_bounds
    for total, world in ((100, 8), (7, 8), (1000000, 3), (0, 2)):
        spans = [shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from hilbert_quantization_b200.distributed import allgather_merge, shard_bounds, merge_topk_host
from oracle import hilbert_oracle as O
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
rng = np.random.default_rng(0)
N, D, Q, k = 400, 64, 6, 5
db = rng.standard_normal((N, D)).astype(np.float32); db[300] = db[10]
qs = rng.standard_normal((Q, D)).astype(np.float32); qs[0] = db[10]
lo, hi = shard_bounds(N, 2, rank)
ids = np.full((Q, k), -1, dtype=np.int64); sc = np.full((Q, k), -1, dtype=np.float32)
for j in range(Q):                        # local exact top-k stands in for the device shard
    i, s = O.topk_stable(np.arange(lo, hi), O.cosine01(qs[j], db[lo:hi]), k)
    ids[j, :len(i)] = i; sc[j, :len(i)] = s.astype(np.float32)
gi, gs = allgather_merge(torch.from_numpy(ids), torch.from_numpy(sc), k, merge_on_host=True)
for j in range(Q):
    i, s = O.topk_stable(np.arange(N), O.cosine01(qs[j], db).astype(np.float32), k)
    assert list(gi[j].numpy()) == list(i), (rank, j, gi[j], i)
assert list(gi[0][:2].numpy()) == [10, 300]
# packed results (what search_batch returns): ids and scores are views of one buffer -> ONE all-gather
from hilbert_quantization_b200.search import packed_result_buffers
pi, ps = packed_result_buffers(Q, k, "cpu")
pi.copy_(torch.from_numpy(ids)); ps.copy_(torch.from_numpy(sc))
calls = []
real = dist.all_gather_into_tensor
dist.all_gather_into_tensor = lambda *a, **kw: (calls.append(1), real(*a, **kw))[1]
gi2, gs2 = allgather_merge(pi, ps, k, merge_on_host=True)
dist.all_gather_into_tensor = real
assert len(calls) == 1, calls
assert torch.equal(gi2, gi) and torch.equal(gs2, gs)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_merge_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


_WORKER_GLOBAL_CUT = r'''
import sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from hilbert_quantization_b200.distributed import global_ratio_cut, shard_bounds
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
rng = np.random.default_rng(7)
Q, N = 9, 1000
scores = rng.random((Q, N)).astype(np.float32)
scores[:, 700] = scores[:, 3]; scores[:, 701] = scores[:, 3]; scores[:, 40] = scores[:, 3]     # exact ties across shards
scores[4] = np.float32(0.75)                                                                    # one query: everything tied
alive = rng.random((Q, N)) < 0.8
thr, ratio = 0.35, 0.3
passed = alive & (scores >= thr)
passed[7] = False; passed[7, 5] = True                                                           # fewer passes than the cap
lo, hi = shard_bounds(N, 2, rank)
if len(sys.argv) > 4 and sys.argv[4] == "empty":                       # rank 1 owns no rows: it must still take part in every collective
    lo, hi = (0, N) if rank == 0 else (N, N)
keep, n_out = global_ratio_cut(torch.from_numpy(scores[:, lo:hi].copy()), torch.from_numpy(passed[:, lo:hi].copy()),
                               torch.from_numpy(alive[:, lo:hi].sum(1)), ratio, lo)
for q in range(Q):
    cap = max(1, int(alive[q].sum() * ratio))
    ids = np.nonzero(passed[q])[0]
    order = ids[np.lexsort((ids, -scores[q, ids]))][:cap]             # score desc, ties -> lower global id
    want = np.zeros(N, bool); want[order] = True
    assert np.array_equal(keep[q].numpy(), want[lo:hi]), (rank, q)
    assert int(n_out[q]) == len(order), (rank, q, int(n_out[q]), len(order))
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_global_ratio_cut_with_an_empty_shard_over_gloo(tmp_path):
    script = tmp_path / "worker_cut_empty.py"
    script.write_text(_WORKER_GLOBAL_CUT)
    port = str(33500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), "empty"], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


def test_two_rank_global_ratio_cut_over_gloo(tmp_path):
    """filter_scope="global": the distributed exact selection equals the single-list cut."""
    script = tmp_path / "worker_cut.py"
    script.write_text(_WORKER_GLOBAL_CUT)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


def test_ingest_plan_codes_describe_run_means_of_the_input_row():
    """host logic of the one-pass shard ingest (search.ingest_plan_codes): every slot of the variant-C compact index row is
    the mean of an aligned run of 4^k consecutive stream values, in exactly the order the oracle's block means over the
    Hilbert grid produce (rag/embedding_generation/hierarchical_index_generator.py:138-178)."""
    from hilbert_quantization_b200.search import ingest_plan_codes
    from oracle import hilbert_oracle as O
    assert ingest_plan_codes(8) is None and ingest_plan_codes(16) is None          # levels below 3: the fused item kernel's job
    rng = np.random.default_rng(5)
    for n, D in ((32, 768), (64, 1536), (64, 4096)):
        codes = ingest_plan_codes(n)
        x = rng.standard_normal((4, D)).astype(np.float32)
        want = O.index_c_batch_compact(O.map_to_2d_batch(x, n))
        assert codes is not None and len(codes) == want.shape[1]
        pad = np.zeros((4, n * n))
        pad[:, :D] = x
        got = np.stack([pad[:, (c & 0xFFFFFF) * 4 ** (c >> 24):((c & 0xFFFFFF) + 1) * 4 ** (c >> 24)].mean(axis=1) for c in codes.tolist()], axis=1)
        assert np.abs(got - want).max() < 3e-7


def test_embedding_frame_dataclass_validates_like_the_reference():
    """frames.EmbeddingFrame mirrors rag/models.py:38-59 (same fields, same checks); no GPU needed."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("hq_frames_src", os.path.join(ROOT, "hilbert_quantization_b200", "frames.py"))
    src = open(spec.origin).read()
    assert "class EmbeddingFrameBatch" in src and "class QuantizedModelBatch" in src
    from hilbert_quantization_b200.frames import EmbeddingFrame
    ok = EmbeddingFrame(np.zeros((67, 64), np.float32), [], 1536, (64, 64), 0.8, 0)
    assert ok.hilbert_dimensions == (64, 64)
    for bad in (dict(original_embedding_dimensions=0), dict(hilbert_dimensions=(64,)), dict(compression_quality=1.5),
                dict(frame_number=-1), dict(embedding_data=np.zeros(3, np.float32))):
        kw = dict(embedding_data=np.zeros((67, 64), np.float32), hierarchical_indices=[], original_embedding_dimensions=1536,
                  hilbert_dimensions=(64, 64), compression_quality=0.8, frame_number=0)
        kw.update(bad)
        with pytest.raises(ValueError):
            EmbeddingFrame(**kw)
