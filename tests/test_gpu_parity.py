"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the fixtures
frozen from the real reference.  Bit-exact for coordinates, permutations, uint8 frames and
variant-B indices; float tolerances are written next to each assertion."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import hilbert_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hq():
    import hilbert_quantization_b200 as m
    return m


def dev_t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------ a1/a2 coordinates
@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 64, 128])
def test_coordinates_bit_exact(hq, n):
    m = hq.HilbertCurveMapper()
    got = m.generate_hilbert_coordinates(n)
    assert isinstance(got, list) and isinstance(got[0], tuple)
    want = load_golden("coords.npz")[f"xy_n{n}"]
    assert got == [tuple(r) for r in want.tolist()]
    assert hq.HilbertCurveMapperImpl(None).generate_hilbert_coordinates(n) == got


def test_coordinates_large_and_inverse(hq):
    m = hq.HilbertCurveMapper()
    for n in (256, 1024, 4096):
        x, y = m._coords_device(n)
        d = torch.arange(n * n, device="cuda")
        # inverse on device, bit exact
        from hilbert_quantization_b200._lib import lib, check
        from hilbert_quantization_b200 import _device as dv
        out = torch.empty(n * n, dtype=torch.int64, device="cuda")
        check(lib.hq_xy2d_batch(n, dv.ptr(x), dv.ptr(y), n * n, dv.ptr(out), dv.stream_ptr()))
        assert torch.equal(out, d)
        # sample against the oracle
        idx = np.random.default_rng(n).integers(0, n * n, 5000)
        ox, oy = O.d2xy(n, idx)
        assert np.array_equal(x.cpu().numpy()[idx], ox) and np.array_equal(y.cpu().numpy()[idx], oy)
        # consecutive positions are 4-neighbours (curve continuity)
        dx = (x[1:] - x[:-1]).abs() + (y[1:] - y[:-1]).abs()
        assert int(dx.max()) == 1 and int(dx.min()) == 1
    assert m._hilbert_index_to_xy(5, 4) == tuple(int(v[0]) for v in O.d2xy(4, [5]))
    assert m._xy_to_hilbert_index(3, 2, 8) == int(O.xy2d(8, [3], [2])[0])


def test_coordinate_errors(hq):
    with pytest.raises(hq.HilbertQuantizationError, match="must be a power of 2"):
        hq.HilbertCurveMapper().generate_hilbert_coordinates(3)
    with pytest.raises(ValueError, match="must be a power of 2"):
        hq.HilbertCurveMapperImpl(None).generate_hilbert_coordinates(6)


# ------------------------------------------------------------------ a3/a4 map / unmap
def test_map_unmap_golden(hq):
    g = load_golden("maps.npz")
    m, r = hq.HilbertCurveMapper(), hq.HilbertCurveMapperImpl(None)
    for key in [k[3:] for k in g.files if k.startswith("in_")]:
        n = int(key.split("_")[0][1:])
        p = g[f"in_{key}"]
        img = m.map_to_2d(p, (n, n))
        assert img.dtype == p.dtype and np.array_equal(img, g[f"map_to_2d_{key}"])
        assert np.array_equal(r.map_to_2d(p, (n, n)), img)
        back = m.map_from_2d(img)
        assert back.dtype == p.dtype and np.array_equal(back, g[f"map_from_2d_{key}"])
        assert np.array_equal(r.map_from_2d(img), back)
    assert list(m.map_from_2d(g["golden_from_in"])) == [1, 2, 3, 4]
    assert list(r.map_from_2d(g["golden_rag_from_in"])) == [1, 3, 4, 2]


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float16, np.int32, np.float32, np.int64, np.float64])
@pytest.mark.parametrize("n,D", [(4, 16), (8, 37), (32, 768), (64, 1536), (128, 10000)])
def test_map_unmap_dtypes(hq, dtype, n, D):
    rng = np.random.default_rng(D)
    p = (rng.standard_normal((3, D)) * 50).astype(dtype)
    m = hq.HilbertCurveMapper()
    words = np.ascontiguousarray(p).view({1: np.uint8, 2: np.int16, 4: np.int32, 8: np.int64}[p.dtype.itemsize])
    grids = m.map_to_2d_batch(dev_t(words), n)
    got = grids.cpu().numpy().view(dtype)
    assert np.array_equal(got.view(np.uint8), O.map_to_2d_batch(p, n).view(np.uint8))
    back = m.map_from_2d_batch(grids).cpu().numpy().view(dtype)
    assert np.array_equal(back.view(np.uint8), O.map_from_2d_batch(O.map_to_2d_batch(p, n)).view(np.uint8))
    trimmed = m.map_from_2d_batch(grids, length=D).cpu().numpy().view(dtype)
    assert np.array_equal(trimmed.view(np.uint8), p.view(np.uint8))


@pytest.mark.parametrize("n,D,N", [(4, 5, 300), (8, 64, 77), (16, 255, 50), (32, 1024, 1000), (64, 1536, 500),
                                   (64, 4096, 33), (128, 16384, 3), (256, 40001, 2), (512, 262144, 1), (1024, 600000, 1)])
def test_map_batch_fp32_bit_exact(hq, n, D, N):
    rng = np.random.default_rng(n + D)
    p = rng.standard_normal((N, D)).astype(np.float32)
    m = hq.HilbertCurveMapper()
    grids = m.map_to_2d_batch(dev_t(p), n)
    want = O.map_to_2d_batch(p, n)
    assert np.array_equal(grids.cpu().numpy(), want)
    assert np.array_equal(m.map_from_2d_batch(grids).cpu().numpy(), O.map_from_2d_batch(want))
    assert np.array_equal(m.map_from_2d_batch(grids, length=D).cpu().numpy(), p)


def test_map_unaligned_rows(hq):
    # row pitch not a multiple of 4 floats -> scalar access path
    rng = np.random.default_rng(5)
    p = rng.standard_normal((9, 1001)).astype(np.float32)
    m = hq.HilbertCurveMapper()
    grids = m.map_to_2d_batch(dev_t(p), 32)
    assert np.array_equal(grids.cpu().numpy(), O.map_to_2d_batch(p, 32))
    assert np.array_equal(m.map_from_2d_batch(grids, length=1001).cpu().numpy(), p)


def test_round_trip_property_large(hq):
    # SURVEY C3 shape (scaled down): 200k x 1024 -> 32x32 -> back, bitwise
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((200_000, 1024), device="cuda", generator=g)
    m = hq.HilbertCurveMapper()
    back = m.map_from_2d_batch(m.map_to_2d_batch(x, 32))
    assert torch.equal(back, x)


def test_mapper_errors(hq):
    m, r = hq.HilbertCurveMapper(), hq.HilbertCurveMapperImpl(None)
    p = np.arange(5, dtype=np.float32)
    with pytest.raises(hq.HilbertQuantizationError, match="requires square dimensions"):
        m.map_to_2d(p, (4, 8))
    with pytest.raises(hq.HilbertQuantizationError, match="must be a power of 2"):
        m.map_to_2d(p, (3, 3))
    with pytest.raises(hq.HilbertQuantizationError, match="Too many parameters"):
        m.map_to_2d(np.arange(20, dtype=np.float32), (4, 4))
    with pytest.raises(hq.HilbertQuantizationError, match="requires square dimensions"):
        m.map_from_2d(np.zeros((4, 8), dtype=np.float32))
    with pytest.raises(ValueError, match="Dimensions must be positive"):
        r.map_to_2d(p, (0, 0))
    with pytest.raises(ValueError, match="Input must be 2D array"):
        r.map_from_2d(np.zeros(16, dtype=np.float32))
    assert np.array_equal(r.map_to_2d(np.array([], dtype=np.float32), (4, 4)), np.zeros((4, 4), dtype=np.float32))


# ------------------------------------------------------------------ a6/a7/a8 indices
def test_index_golden(hq):
    g = load_golden("index.npz")
    a_gen = hq.HierarchicalIndexGeneratorImpl()
    b_gen = hq.StreamingHilbertIndexGenerator()
    c_gen = hq.HierarchicalIndexGenerator()
    for key in [k[3:] for k in g.files if k.startswith("in_n")]:
        n = int(key.split("_")[0][1:])
        img = O.map_to_2d(g[f"in_{key}"], (n, n))
        b = b_gen.generate_optimized_indices(img, n)
        assert b.dtype == np.float64 and np.array_equal(b, g[f"B_{key}"])              # bit exact
        a = a_gen.generate_optimized_indices(img, n)
        assert a.dtype == np.float32 and np.abs(a - g[f"A_{key}"]).max() <= 3e-7        # fp32 tree vs numpy pairwise
        c = c_gen.generate_multi_level_indices(img)
        assert c.shape == g[f"C_{key}"].shape and c.dtype == np.float32
        assert np.array_equal(c[:n], img) and np.abs(c - g[f"C_{key}"]).max() <= 3e-7
    img = O.map_to_2d(g["in_n128_D16384"], (128, 128))
    assert np.abs(a_gen.generate_optimized_indices(img, 1024) - g["A_S1024_n128"]).max() <= 3e-7
    _, idx, stats = b_gen.generate_indices_during_mapping(g["B_during_mapping_in"], (32, 32), 32)
    assert np.array_equal(idx, g["B_during_mapping_out"]) and stats["total_values_processed"] == 1000


@pytest.mark.parametrize("n,D,N", [(4, 16, 600), (8, 50, 130), (16, 256, 70), (32, 768, 333), (32, 1024, 64),
                                   (64, 1536, 257), (64, 4096, 9), (128, 9000, 5), (256, 65536, 2), (512, 200000, 1)])
def test_fused_map_index_batch(hq, n, D, N):
    rng = np.random.default_rng(n * 7 + D)
    p = rng.standard_normal((N, D)).astype(np.float32)
    grids_want = O.map_to_2d_batch(p, n)
    t = dev_t(p)
    # C compact + grid in one pass
    grids, idx = hq.map_and_index(t, n, variant="C")
    assert np.array_equal(grids.cpu().numpy(), grids_want)
    want = O.index_c_batch_compact(grids_want)
    assert idx.shape == want.shape
    assert np.abs(idx.cpu().numpy() - want).max() <= 3e-7                              # fp32 tree means, N(0,1) data
    # enhanced frames (grid + zero padded index rows)
    frames, _ = hq.map_and_index(t, n, variant="C", enhanced=True)
    fw = np.stack([O.index_c(gw) for gw in grids_want[: min(N, 8)]])
    f = frames.cpu().numpy()
    assert np.array_equal(f[:, :n], grids_want) and np.abs(f[: fw.shape[0]] - fw).max() <= 3e-7
    # B (float64, bit exact) and A
    for S in (n, 37):
        _, b = hq.map_and_index(t, n, variant="B", index_space=S, want_grid=False)
        bw = np.stack([O.index_b(gw, S) for gw in grids_want[: min(N, 6)]])
        assert b.dtype == torch.float64 and np.array_equal(b.cpu().numpy()[: bw.shape[0]], bw)
        _, a = hq.map_and_index(t, n, variant="A", index_space=S, want_grid=False)
        aw = np.stack([O.index_a(gw, S) for gw in grids_want[: min(N, 6)]])
        assert np.abs(a.cpu().numpy()[: aw.shape[0]] - aw).max() <= 3e-7
    # from already-mapped grids (direction 1)
    idx2 = hq.index_from_grids(grids, variant="C")
    assert torch.equal(idx2, idx)


@pytest.mark.parametrize("n,total", [(128, 3 * 128 * 128 + 5000), (256, 65536 * 2), (128, 128 * 128 - 4), (256, 70000),
                                     (128, 2 * 128 * 128 + 4098)])
def test_parameter_stream_ragged_tail(hq, n, total):
    """BASELINE config 4 shape: a flat parameter stream cut into n x n grids, the last one zero padded, ONE launch
    (hq_map_index_stream); total = 2 * 128 * 128 + 4098 is not a multiple of 4 -> two-pass fallback, same results."""
    rng = np.random.default_rng(total)
    stream = (0.02 * rng.standard_normal(total)).astype(np.float32)
    cells = n * n
    N = -(-total // cells)
    padded = np.zeros(N * cells, dtype=np.float32)
    padded[:total] = stream
    grids_want = O.map_to_2d_batch(padded.reshape(N, cells), n)
    t = dev_t(stream)
    grids, idx = hq.map_parameter_stream(t, n, variant="C")
    assert np.array_equal(grids.cpu().numpy(), grids_want)
    want = O.index_c_batch_compact(grids_want)
    assert idx.shape == want.shape and np.abs(idx.cpu().numpy() - want).max() <= 1e-8        # values ~ 0.02: 3e-7 relative
    # same numbers as the per-grid API on the padded stream
    g2, i2 = hq.map_and_index(dev_t(padded.reshape(N, cells)), n, variant="C")
    assert torch.equal(g2, grids) and torch.equal(i2, idx)
    _, b = hq.map_parameter_stream(t, n, variant="B", index_space=n, want_grid=False)
    bw = np.stack([O.index_b(gw, n) for gw in grids_want])
    assert np.array_equal(b.cpu().numpy(), bw)
    back = hq.HilbertCurveMapper().map_from_2d_batch(grids).cpu().numpy().reshape(-1)
    assert np.array_equal(back[:total], stream) and not back[total:].any()


def test_index_c_general_shapes(hq):
    rng = np.random.default_rng(11)
    c_gen = hq.HierarchicalIndexGenerator()
    for shape in ((2, 2), (6, 10), (12, 12), (16, 32), (2, 4)):
        img = rng.standard_normal(shape).astype(np.float32)
        got = c_gen.generate_multi_level_indices(img)
        want = O.index_c(img)
        assert got.shape == want.shape and np.abs(got - want).max() <= 3e-7
    quad = np.zeros((4, 4), dtype=np.float32)
    quad[:2, :2], quad[:2, 2:], quad[2:, 2:], quad[2:, :2] = 2, 3, 4, 5
    assert np.array_equal(c_gen._calculate_hilbert_order_averages(quad, 2), np.array([2, 3, 4, 5], dtype=np.float32))
    ramp = np.arange(16, dtype=np.float32).reshape(4, 4)
    assert np.allclose(hq.HierarchicalIndexGeneratorImpl().calculate_spatial_averages(ramp, 2), [2.5, 4.5, 10.5, 12.5])


# ------------------------------------------------------------------ a10 uint8
def test_quantize_bit_exact(hq):
    g = load_golden("index.npz")
    fq = hq.FrameQuantizer()
    for key in [k[3:] for k in g.files if k.startswith("in_n")]:
        enh = g[f"enhB_{key}"]
        q = fq._normalize_for_compression(enh)
        assert q.dtype == np.uint8 and np.array_equal(q, g[f"u8_{key}"])
        assert np.array_equal(fq._denormalize_from_compression(q), g[f"deq_{key}"])
    const = np.full((5, 4), 2.5, dtype=np.float32)
    assert np.array_equal(hq.FrameQuantizer()._normalize_for_compression(const), np.full((5, 4), 128, dtype=np.uint8))


def test_quantize_batch_and_large(hq):
    rng = np.random.default_rng(2)
    frames = (rng.standard_normal((300, 65, 64)) * 3).astype(np.float32)
    frames[7] = 1.0
    q, mm = hq.quantize_u8_batch(dev_t(frames))
    qn, mmn = q.cpu().numpy(), mm.cpu().numpy()
    for i in (0, 7, 150, 299):
        want, mn, mx = O.normalize_u8(frames[i])
        assert np.array_equal(qn[i], want) and mmn[i, 0] == mn and mmn[i, 1] == mx
    deq = hq.dequantize_u8_batch(q, mm).cpu().numpy()
    for i in (0, 7, 299):
        _, mn, mx = O.normalize_u8(frames[i])
        assert np.array_equal(deq[i], O.denormalize_u8(qn[i], mn, mx))
    big = (rng.standard_normal((2, 513, 512)) * 2).astype(np.float32)       # multi-block path
    q, mm = hq.quantize_u8_batch(dev_t(big))
    for i in range(2):
        want, mn, mx = O.normalize_u8(big[i])
        assert np.array_equal(q[i].cpu().numpy(), want) and float(mm[i, 0]) == mn and float(mm[i, 1]) == mx


def test_fused_map_index_quantize_matches_the_reference_pipeline(hq):
    """North star (2): map_to_2d + index + embed + uint8 normalise in ONE kernel.  The golden `u8_*` frames are the
    REFERENCE's _normalize_for_compression(embed_indices_in_image(map_to_2d(p), B-indices)) (core/pipeline.py:129-146,
    core/compressor.py:256-280): bit-exact, including the (min, max) pair.  32 x 32 and 64 x 64 run the fused kernel, the
    other sizes the two-launch path."""
    g = load_golden("index.npz")
    for key in [k[3:] for k in g.files if k.startswith("in_n")]:
        n = int(key[1:].split("_")[0])
        p = g[f"in_{key}"]
        frames, mm, idx = hq.map_index_quantize(dev_t(p[None, :]), n, variant="B", want_indices=True)
        want = g[f"u8_{key}"]
        assert frames.dtype == torch.uint8 and tuple(frames.shape) == (1, n + 1, n), key
        assert np.array_equal(frames[0].cpu().numpy(), want), (key, int((frames[0].cpu().numpy() != want).sum()))
        enh = g[f"enhB_{key}"]
        assert float(mm[0, 0]) == float(enh.min()) and float(mm[0, 1]) == float(enh.max()), key
        assert np.array_equal(idx[0].cpu().numpy(), g[f"B_{key}"]), key


@pytest.mark.parametrize("n,D,variant", [(64, 1536, "C"), (64, 4096, "B"), (64, 1536, "A"), (32, 768, "C"), (32, 1024, "A"),
                                         (32, 768, "B"), (64, 1000, "B"), (32, 12, "C")])
def test_fused_map_index_quantize_batches(hq, n, D, variant):
    """Batches (partial last chunk, constant rows, all-positive rows whose padding pulls the minimum to 0) against the oracle:
    frame = grid + index rows in float32, then O.normalize_u8; and against the two-launch device path."""
    rng = np.random.default_rng(n * D)
    N = 1037
    x = (rng.standard_normal((N, D)) * 2).astype(np.float32)
    x[3] = 1.5                                            # constant vector (a constant FRAME only when D == n * n and the index is constant)
    x[4] = np.abs(x[4]) + 0.25                            # all positive: min comes from the zero padding, if any
    x[5] = -np.abs(x[5]) - 0.25                           # all negative: max comes from the zero padding, if any
    x[6] = 0.0
    frames, mm, idx = hq.map_index_quantize(dev_t(x), n, variant=variant, want_indices=True)
    fr, mmn = frames.cpu().numpy(), mm.cpu().numpy()
    for i in (0, 3, 4, 5, 6, 511, 1024, N - 1):
        img = O.map_to_2d(x[i], (n, n))
        if variant == "C":
            enh = O.index_c(img)
        else:
            row = O.index_a(img, n) if variant == "A" else O.index_b(img, n)
            enh = O.embed_index_row(img, row)
        if variant == "B":
            want, mn, mx = O.normalize_u8(enh.astype(np.float32))
            assert np.array_equal(fr[i], want), (i, int((fr[i] != want).sum()))
            assert mmn[i, 0] == mn and mmn[i, 1] == mx, i
        else:       # fp32 block means: tree vs pairwise order (<= 3e-7 on the index values): a byte may differ where a value sits on a bin edge
            want, mn, mx = O.normalize_u8(enh.astype(np.float32))
            assert np.array_equal(fr[i, :n], want[:n]), i                       # the grid part is exact
            assert mmn[i, 0] == mn and mmn[i, 1] == mx, i
            assert np.abs(fr[i, n:].astype(int) - want[n:].astype(int)).max() <= 1, i
    # the same frames from the two-launch device path (fused map + index, then hq_quantize_u8): bit-identical
    if variant == "C":
        enh_dev, _ = hq.map_and_index(dev_t(x), n, variant="C", enhanced=True)
    else:
        grid, ix = hq.map_and_index(dev_t(x), n, variant=variant, index_space=n)
        enh_dev = torch.cat([grid, ix.to(torch.float32).unsqueeze(1)], dim=1)
        assert torch.equal(ix, idx)
    q2, mm2 = hq.quantize_u8_batch(enh_dev)
    assert torch.equal(q2, frames) and torch.equal(mm2, mm)


def test_fused_quantize_reciprocal_division_is_exact_over_many_scales(hq):
    """The fused kernel replaces the per-cell IEEE division by the item's reciprocal + one exact-residual correction
    (hq_item_pass.cuh, quant_u8).  80 M cells at scales from 1e-25 to 1e25 (and beyond, where it falls back to the division):
    every byte must equal hq_quantize_u8's, which divides."""
    g = torch.Generator(device="cuda").manual_seed(99)
    N, D, n = 20000, 4096, 64
    x = torch.randn((N, D), generator=g, device="cuda")
    scale = torch.pow(10.0, torch.rand((N, 1), generator=g, device="cuda") * 70.0 - 35.0)
    x = (x * scale + torch.randn((N, 1), generator=g, device="cuda") * scale * 3).contiguous()
    frames, mm, _ = hq.map_index_quantize(x, n, variant="C")
    enh, _ = hq.map_and_index(x, n, variant="C", enhanced=True)
    q2, mm2 = hq.quantize_u8_batch(enh)
    assert torch.equal(mm, mm2)
    assert torch.equal(frames, q2), int((frames != q2).sum())


# ------------------------------------------------------------------ a12/a13/a15 search
def _oracle_rows(db, n):
    levels = O.c_granularity_levels(n)
    compact = O.index_c_batch_compact(O.map_to_2d_batch(db, n))
    rows, o = [], 0
    for gr in levels:
        r = np.zeros((db.shape[0], n), dtype=np.float32)
        r[:, : gr * gr] = compact[:, o:o + gr * gr]
        rows.append(r)
        o += gr * gr
    return rows


def _check_query(hq, db, rows, q, n, k, ids, scores, mask_row):
    """Compare one query's device result with the oracle, tolerating only borderline cases."""
    from hilbert_quantization_b200 import search as S
    qc = O.index_c_batch_compact(O.map_to_2d_batch(q[None], n))[0]
    q_rows, o = [], 0
    for gr in O.c_granularity_levels(n):
        r = np.zeros(n, dtype=np.float32)
        r[: gr * gr] = qc[o:o + gr * gr]
        q_rows.append(r)
        o += gr * gr
    surv, trace = O.rag_progressive_filter(q_rows, rows, return_scores=True)
    got = set(np.nonzero(mask_row)[0].tolist())
    diff = got ^ set(surv.tolist())
    if diff:
        # a differing row must sit within 2e-6 of a threshold or of the rank-cut score at some level
        border = set()
        for cand, s, thr, cap in trace:
            border |= set(cand[np.abs(s - thr) < 2e-6].tolist())
            if len(s) > cap:
                cut = np.sort(s)[::-1][cap - 1]
                border |= set(cand[np.abs(s - cut) < 2e-6].tolist())
        assert diff <= border, f"{len(diff)} non-borderline filter differences"
        return False
    ids_w, sc_w = O.progressive_search(q, db, n, k, db_rows=rows)
    m = len(ids_w)
    assert list(ids[:m]) == list(ids_w) and all(i == -1 for i in ids[m:])
    assert np.abs(scores[:m] - sc_w).max() < 5e-7                      # (cos+1)/2 in fp32 vs fp64 oracle
    return True


def test_search_golden(hq):
    g = load_golden("rag_search.npz")
    for tag in "abc":
        db, qs, n = g[f"{tag}_db"], g[f"{tag}_queries"], int(g[f"{tag}_n"])
        d = hq.EmbeddingDatabase(db, n=n)
        ids, scores, mask = hq.search_batch(d, qs, 10, return_mask=True)
        from hilbert_quantization_b200.search import unpack_mask
        alive = unpack_mask(mask, db.shape[0])
        ids, scores = ids.cpu().numpy(), scores.cpu().numpy()
        for j in range(len(qs)):
            want = g[f"{tag}_survivors_q{j}"]
            assert set(np.nonzero(alive[j])[0].tolist()) == set(want.tolist())
            kk = len(g[f"{tag}_topk_ids_q{j}"])
            assert list(ids[j, :kk]) == list(g[f"{tag}_topk_ids_q{j}"])
            assert np.abs(scores[j, :kk] - g[f"{tag}_topk_scores_q{j}"]).max() < 5e-7


@pytest.mark.parametrize("n,D,N,Q", [(16, 256, 3000, 12), (32, 768, 5000, 20), (32, 1024, 4097, 9), (64, 1536, 6000, 16)])
def test_search_vs_oracle(hq, n, D, N, Q):
    from hilbert_quantization_b200.search import unpack_mask
    rng = np.random.default_rng(N + D)
    db = rng.standard_normal((N, D)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    db[100] = db[50]                                                   # duplicates: tie -> lower id
    db[N - 1] = db[50]
    qs = []
    for j in range(Q):
        # near-duplicate of a stored row (noise norm 0.3) or a fresh random direction
        q = db[(j * 37) % N] + (0.3 / np.sqrt(D)) * rng.standard_normal(D).astype(np.float32) if j % 2 == 0 else rng.standard_normal(D).astype(np.float32)
        qs.append((q / np.linalg.norm(q)).astype(np.float32))
    qs = np.stack(qs)
    qs[2] = db[50]
    d = hq.EmbeddingDatabase(db, n=n)
    ids, scores, mask = hq.search_batch(d, qs, 10, return_mask=True)
    alive = unpack_mask(mask, N)
    ids, scores = ids.cpu().numpy(), scores.cpu().numpy()
    rows = _oracle_rows(db, n)
    exact = sum(_check_query(hq, db, rows, qs[j], n, 10, ids[j], scores[j], alive[j]) for j in range(Q))
    assert exact >= Q - 1                                              # borderline survivor sets are rare
    assert ids[2, 0] == 50 and list(ids[2, :3]) == [50, 100, N - 1]    # exact ties resolve to the lower id
    for j in range(0, Q, 2):
        if j != 2:
            assert ids[j, 0] == (j * 37) % N                           # perturbed rows find their source


def test_search_chunked_and_unfiltered(hq):
    rng = np.random.default_rng(9)
    db = rng.standard_normal((2000, 768)).astype(np.float32)
    qs = rng.standard_normal((7, 768)).astype(np.float32)
    d = hq.EmbeddingDatabase(db)
    a = hq.search_batch(d, qs, 5)
    b = hq.search_batch(d, qs, 5, work_bytes=3 * 4 * 2000)             # forces 3-query chunks
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    ids, scores = hq.search_batch(d, qs, 5, use_filter=False)
    for j in range(7):
        iw, sw = O.topk_stable(np.arange(2000), O.cosine01(qs[j], db), 5)
        assert list(ids[j].cpu().numpy()) == list(iw) and np.abs(scores[j].cpu().numpy() - sw).max() < 5e-7


def test_filter_ratio_cut_binds(hq):
    # all-positive data: every row passes every threshold, so the ratio cut (exact radix select) decides
    from hilbert_quantization_b200.search import unpack_mask
    rng = np.random.default_rng(4)
    N, D, n = 5000, 768, 32
    db = (rng.random((N, D)) + 0.5).astype(np.float32)
    db[10] = db[20]
    qs = (rng.random((5, D)) + 0.5).astype(np.float32)
    d = hq.EmbeddingDatabase(db, n=n)
    ids, scores, mask = hq.search_batch(d, qs, 10, return_mask=True)
    alive = unpack_mask(mask, N)
    rows = _oracle_rows(db, n)
    for j in range(5):
        assert alive[j].sum() == max(1, int(max(1, int(N * 0.3)) * 0.5))
        _check_query(hq, db, rows, qs[j], n, 10, ids[j].cpu().numpy(), scores[j].cpu().numpy(), alive[j])


def test_rag_engine_surface(hq):
    rng = np.random.default_rng(21)
    n, D, N = 32, 768, 150
    db = rng.standard_normal((N, D)).astype(np.float32)
    frames = [O.index_c(O.map_to_2d(v, (n, n))) for v in db]
    q = db[5] + 0.05 * rng.standard_normal(D).astype(np.float32)
    qf = O.index_c(O.map_to_2d(q.astype(np.float32), (n, n)))
    eng = hq.RAGSearchEngineImpl(None)
    eng._get_all_candidate_embeddings = lambda: frames
    got = eng.progressive_hierarchical_search(qf)
    L = len(O.c_granularity_levels(n))
    want = O.rag_progressive_filter([qf[n + l] for l in range(L)], [np.stack([f[n + l] for f in frames]) for l in range(L)])
    assert got == list(want)
    assert abs(eng._calculate_embedding_cosine_similarity(qf[:n], frames[5][:n]) - O.cosine01(q, db[5:6])[0]) < 5e-7
    assert eng.progressive_hierarchical_search(np.array([])) == []


def test_core_search_golden(hq):
    g = load_golden("core_search.npz")

    class M:
        def __init__(self, idx, i):
            self.hierarchical_indices, self.i = idx, i
    for S in (32, 64):
        c, q = g[f"S{S}_cands"], g[f"S{S}_query"]
        eng = hq.ProgressiveSimilaritySearchEngine(similarity_threshold=0.1, max_candidates_per_level=20)
        sims = eng._level_sims(q, list(c))
        assert np.abs(sims - g[f"S{S}_level_sims"]).max() < 1e-12
        res = eng.progressive_search(q, [M(c[i], i) for i in range(len(c))], 10)
        assert [r.model.i for r in res] == list(g[f"S{S}_ids"])
        assert np.abs(np.array([r.similarity_score for r in res]) - g[f"S{S}_scores"]).max() < 1e-12
        assert abs(eng.compare_indices_at_level(q, c[3], 1) - g[f"S{S}_level_sims"][3, 1]) < 1e-12


def test_topk_merge_kernel(hq):
    from hilbert_quantization_b200._lib import lib, check
    from hilbert_quantization_b200 import _device as dv
    from hilbert_quantization_b200.distributed import merge_topk_host
    rng = np.random.default_rng(8)
    P, Q, k = 8, 50, 10
    ids = rng.permutation(P * Q * k).reshape(P, Q, k).astype(np.int64)
    sc = np.round(rng.random((P, Q, k)), 2).astype(np.float32)          # many exact ties
    ids[3, :, 7:] = -1
    ti, ts = dev_t(ids), dev_t(sc)
    oi = torch.empty((Q, k), dtype=torch.int64, device="cuda")
    os_ = torch.empty((Q, k), dtype=torch.float32, device="cuda")
    check(lib.hq_topk_merge(dv.ptr(ti), dv.ptr(ts), P, Q, k, dv.ptr(oi), dv.ptr(os_), dv.stream_ptr()))
    wi, ws = merge_topk_host(ids, sc, k)
    assert np.array_equal(oi.cpu().numpy(), wi) and np.array_equal(os_.cpu().numpy(), ws)
    # packed layout of the single all-gather: per shard [Q*k int64 ids | Q*k float32 scores]
    n = Q * k
    blocks = np.zeros((P, 3 * n), dtype=np.int32)
    blocks[:, : 2 * n] = ids.reshape(P, n).view(np.int32)
    blocks[:, 2 * n:] = sc.reshape(P, n).view(np.int32)
    tb = dev_t(blocks)
    oi.fill_(0); os_.fill_(0)
    check(lib.hq_topk_merge_strided(tb.data_ptr(), tb.data_ptr() + 8 * n, P, Q, k, 3 * n // 2, 3 * n, dv.ptr(oi), dv.ptr(os_),
                                    dv.stream_ptr()))
    assert np.array_equal(oi.cpu().numpy(), wi) and np.array_equal(os_.cpu().numpy(), ws)


@pytest.mark.parametrize("N,Q,k", [(100003, 3, 10), (50000, 1, 100), (9000, 2, 10), (300000, 4, 7)])
def test_topk_chunked_equals_single_level(hq, N, Q, k):
    """Two-level top-k (chunk lists + merge) for few queries over long rows == the one-CTA-per-row kernel, ties included."""
    from hilbert_quantization_b200._lib import lib, check
    from hilbert_quantization_b200 import _device as dv
    rng = np.random.default_rng(N + k)
    sc = np.round(rng.random((Q, N)), 3).astype(np.float32)             # ~1000 distinct values: many exact ties
    sc[rng.random((Q, N)) < 0.7] = -1.0                                  # dead rows
    sc[1 % Q, :] = -1.0
    sc[1 % Q, N - 3:] = 0.5                                              # fewer than k live entries, all in the last chunk
    t = dev_t(sc)
    i1 = torch.empty((Q, k), dtype=torch.int64, device="cuda"); s1 = torch.empty((Q, k), dtype=torch.float32, device="cuda")
    i2 = torch.empty_like(i1); s2 = torch.empty_like(s1)
    check(lib.hq_topk_from_scores(dv.ptr(t), t.stride(0), N, Q, k, 1000, dv.ptr(i1), dv.ptr(s1), dv.stream_ptr()))
    nb = int(lib.hq_topk_chunked_scratch_bytes(N, Q, k))
    scratch = torch.empty(max(nb, 8), dtype=torch.uint8, device="cuda")
    check(lib.hq_topk_from_scores_chunked(dv.ptr(t), t.stride(0), N, Q, k, 1000, dv.ptr(i2), dv.ptr(s2), dv.ptr(scratch), nb,
                                          dv.stream_ptr()))
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    for q in range(Q):
        live = np.nonzero(sc[q] >= 0)[0]
        order = live[np.lexsort((live, -sc[q, live]))][:k]
        assert list(i1[q, : len(order)].cpu().numpy() - 1000) == list(order)
        assert (i1[q, len(order):] == -1).all()


def test_readme_surface(hq):
    rng = np.random.default_rng(1)
    eng = hq.ProgressiveSearchEngine(use_frame_caching=True)
    vecs = rng.standard_normal((200, 384)).astype(np.float32)
    for i, v in enumerate(vecs):
        eng.add_document({"document_id": f"doc{i}", "content": f"text {i}", "embedding": v})
    res = eng.search(vecs[17] + 0.01 * rng.standard_normal(384).astype(np.float32), max_results=5)
    assert res and res[0].document_id == "doc17" and 0.0 <= res[0].similarity_score <= 1.0001
    rag = hq.RAGSystem(embedding_dimension=256)
    rag.add_document("a", "hilbert curves preserve locality of reference")
    rag.add_document("b", "completely different words about cooking pasta")
    out = rag.search("locality of hilbert curves", max_results=2)
    assert out and out[0].document_id == "a" and out[0].content.startswith("hilbert")


# ---------------------------------------------------------------------------------------
# a14 comprehensive similarity blend (rag/search/engine.py:516-575)
# ---------------------------------------------------------------------------------------
def test_comprehensive_blend_golden_and_oracle(hq):
    g = load_golden("rag_blend.npz")
    for tag in ("n64_D1536", "n32_D768", "n16_D200", "n64_D4096"):
        qf, cfs = g[f"{tag}_query_frame"], g[f"{tag}_cand_frames"]
        got = hq.comprehensive_scores(cfs, qf).cpu().numpy()[0]
        assert np.abs(got - g[f"{tag}_comprehensive"]).max() < 2e-6          # fp32 sums vs the reference's NumPy
        assert np.abs(got - O.comprehensive_similarity(qf, cfs, qf.shape[1])).max() < 2e-6
        eng = hq.RAGSearchEngineImpl()
        ranked = eng.calculate_embedding_similarity(qf, {i: c for i, c in enumerate(cfs)})
        assert [r[0] for r in ranked[:5]] == list(g[f"{tag}_ranked_ids"][:5])
        assert abs(eng._calculate_comprehensive_similarity(qf, None, cfs[3]) - g[f"{tag}_comprehensive"][3]) < 2e-6


def test_comprehensive_blend_batch_and_shortlist(hq):
    rng = np.random.default_rng(11)
    N, D, Q, n = 300, 1536, 5, 64
    db = rng.standard_normal((N, D)).astype(np.float32)
    qs = rng.standard_normal((Q, D)).astype(np.float32)
    frames, _ = hq.map_and_index(torch.from_numpy(db).cuda(), n, variant="C", enhanced=True)
    qframes, _ = hq.map_and_index(torch.from_numpy(qs).cuda(), n, variant="C", enhanced=True)
    full = hq.comprehensive_scores(frames, qframes).cpu().numpy()
    fr, qfr = frames.cpu().numpy(), qframes.cpu().numpy()
    for j in range(Q):
        assert np.abs(full[j] - O.comprehensive_similarity(qfr[j], fr, n)).max() < 2e-6
    ids = rng.integers(0, N, size=(Q, 7))
    ids[2, 4] = -1
    short = hq.comprehensive_scores(frames, qframes, cand_ids=ids).cpu().numpy()
    for j in range(Q):
        for m in range(7):
            assert short[j, m] == (-1.0 if ids[j, m] < 0 else full[j, ids[j, m]])


# ---------------------------------------------------------------------------------------
# f1 PrecomputedHilbertIndexer (core/precomputed_hilbert_index.py:65-212)
# ---------------------------------------------------------------------------------------
def test_precomputed_indexer_golden(hq):
    g = load_golden("precomputed.npz")
    for n in (8, 16, 64, 128):
        img = g[f"n{n}_image"]
        ix = hq.PrecomputedHilbertIndexer()
        idx = ix.create_precomputed_index(img, f"m{n}")
        assert ix.get_index(f"m{n}") is idx and idx.original_shape == (n, n)
        want_levels = [tuple(r) for r in g[f"n{n}_levels"]]
        assert [(l.grid_size, l.square_size, l.num_squares) for l in idx.levels] == want_levels
        for i, l in enumerate(idx.levels):
            assert l.averages.dtype == np.float32
            assert np.abs(l.averages - g[f"n{n}_avg{i}"]).max() < 3e-7          # 4-ary tree vs NumPy's pairwise float32 mean
            assert np.array_equal(np.array(l.square_coordinates).reshape(-1, 2), g[f"n{n}_xy{i}"])
    with pytest.raises(ValueError, match="must be square"):
        hq.PrecomputedHilbertIndexer().create_precomputed_index(np.zeros((4, 8), np.float32), "x")


def test_precomputed_indexer_batch_vs_oracle(hq):
    rng = np.random.default_rng(3)
    grids = rng.standard_normal((5, 32, 32)).astype(np.float32)
    ix = hq.PrecomputedHilbertIndexer()
    per_level = ix.level_averages_batch(torch.from_numpy(grids).cuda())
    for (gs, ss), t in zip(O.precomputed_granularity_levels(32), per_level):
        got = t.cpu().numpy()
        for i in range(5):
            assert np.abs(got[i] - O.precomputed_level_averages(grids[i], gs, ss)).max() < 3e-7


# ---------------------------------------------------------------------------------------
# f4 video-path hierarchical similarity / frame ordering (core/video_storage.py)
# ---------------------------------------------------------------------------------------
def test_video_path_similarity_and_ordering_golden(hq):
    g = load_golden("video_order.npz")
    for tag in ("S64_float32", "S1024_float32", "S32_float64"):
        frames, q = list(g[f"{tag}_frames"]), g[f"{tag}_query"]
        sims = hq.video.hierarchical_similarity_matrix(q, np.stack(frames)).cpu().numpy()[0]
        assert np.abs(sims - g[f"{tag}_sims"]).max() < 1e-12
        pair = hq.video.hierarchical_similarity_matrix(np.stack(frames[:12]), np.stack(frames[:12])).cpu().numpy()
        assert np.abs(pair - g[f"{tag}_pair_sims"]).max() < 1e-12
        assert pair[11, 11] == 1.0 and pair[11, 0] == 0.0 and pair[3, 7] == 1.0       # zero variance / duplicate rows
        assert hq.video.sort_frames_by_hierarchical_indices(frames) == list(g[f"{tag}_order"])
        assert [hq.video.find_optimal_insertion_position(q, frames[:15]),
                hq.video.find_optimal_insertion_position(frames[20], frames[:15])] == list(g[f"{tag}_insert_pos"])
        top = hq.video.traditional_search(q, frames, 5)
        want = np.argsort(-g[f"{tag}_sims"], kind="stable")[:5]
        assert [t[0] for t in top] == list(want)
        assert abs(hq.video.calculate_hierarchical_similarity(q, frames[0]) - g[f"{tag}_sims"][0]) < 1e-12


def test_video_hierarchical_search_golden(hq):
    """core/video_search.py:215-262: every frame scored by the core engine's finest-level comparison in one launch;
    ids (ties and skipped frames included) identical to the reference's, scores <= 1e-12."""
    g = load_golden("video_search.npz")
    for tag in ("S64", "S340", "S21"):
        frames = [None if i in set(g[f"{tag}_missing"].tolist()) else f for i, f in enumerate(g[f"{tag}_frames"])]
        res = hq.video.hierarchical_search(g[f"{tag}_query"], frames, 12, float(g[f"{tag}_threshold"]))
        assert [r[0] for r in res] == list(g[f"{tag}_ids"])
        assert np.abs(np.array([r[1] for r in res]) - g[f"{tag}_scores"]).max() < 1e-12
    assert hq.video.hierarchical_search(g["S64_query"], [None, None], 5) == []
    assert hq.video.hierarchical_search(np.zeros(0), list(g["S64_frames"][:4]), 5) == []


# ---------------------------------------------------------------------------------------
# shard ingest: index rows + norms + bf16 unit rows in one pass (hq_shard_ingest)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,D", [(1000, 1536), (777, 768), (513, 1024), (100, 512), (64, 2048), (33, 3072), (9, 4096), (1, 1536)])
def test_shard_ingest_is_bit_identical_to_the_separate_passes(hq, N, D):
    """run means of the input row == block means over the Hilbert grid: the fused pass must reproduce hq_map_index (no grid),
    hq_row_norms and hq_to_bf16_unit bit for bit, and the oracle's index rows within the fp32 bound of test_fused_map_index_batch"""
    from hilbert_quantization_b200.index import map_and_index
    from hilbert_quantization_b200.search import row_norms, shard_ingest, to_bf16
    from hilbert_quantization_b200.dimension import rag_optimal_dimensions
    g = torch.Generator(device="cuda").manual_seed(N + D)
    x = torch.randn((N, D), device="cuda", generator=g)
    if N > 4:
        x[3] = 0.0
        x[4, D // 2:] = 0.0
    n = rag_optimal_dimensions(D)[0]
    fused = shard_ingest(x, n)
    assert fused is not None
    idx, norms, unit = fused
    _, want_idx = map_and_index(x, n, variant="C", layout="compact", want_grid=False)
    want_norms = row_norms(x)
    assert torch.equal(idx.view(torch.int32), want_idx.view(torch.int32))
    assert torch.equal(norms.view(torch.int32), want_norms.view(torch.int32))
    assert torch.equal(unit.view(torch.int16), to_bf16(x, want_norms).view(torch.int16))
    m = min(N, 50)
    ref = O.index_c_batch_compact(O.map_to_2d_batch(x[:m].cpu().numpy(), n))
    assert np.abs(idx[:m].cpu().numpy().astype(np.float64) - ref).max() < 3e-7
    # a row-pitched view of a wider matrix, no bf16 rows
    wide = torch.randn((N, D + 128), device="cuda", generator=g)
    i2, n2, u2 = shard_ingest(wide[:, :D], n, want_bf16=False)
    _, w2 = map_and_index(wide[:, :D].contiguous(), n, variant="C", layout="compact", want_grid=False)
    assert u2 is None and torch.equal(i2, w2) and torch.equal(n2, row_norms(wide[:, :D]))


def test_shard_ingest_declines_shapes_it_does_not_cover(hq):
    from hilbert_quantization_b200.dimension import rag_optimal_dimensions
    from hilbert_quantization_b200.search import shard_ingest
    for D in (250, 256, 640, 1537):                                            # odd widths, 16 x 16 grids (levels below 3)
        assert shard_ingest(torch.randn((10, D), device="cuda"), rag_optimal_dimensions(D)[0]) is None
    d = hq.EmbeddingDatabase(torch.randn((300, 250), device="cuda"))          # falls back to the separate passes
    assert d.idx.shape[0] == 300 and d.emb_bf16 is not None


# ------------------------------------------------------------------ a12: the reference's tie rule at the ratio cuts
def test_exact_ties_at_a_ratio_cut_follow_the_previous_level_like_the_reference(hq):
    """Rows that differ only by a permutation of four level-0 blocks inside ONE level-1 block have identical level-1 and
    level-2 index rows (integer values: every block mean is exact), so their level-1 / level-2 scores tie EXACTLY while their
    level-0 scores differ.  The reference's stable sort (rag/search/engine.py:236) keeps such rows in the previous level's
    order; `tie_rule="reference"` reproduces it (ids equal the oracle's), the default rule (lower row id) provably differs
    on this input -- the deviation DESIGN.md section 8 documents."""
    import itertools
    D, n = 1536, 64
    base = np.full(24, 2.0, dtype=np.float32)
    rows = []
    for perm in itertools.permutations([1.0, 2.0, 3.0, 5.0]):
        b = base.copy()
        b[:4] = perm
        rows.append(np.repeat(b, 64))
    db = np.stack(rows).astype(np.float32)                      # 24 rows, constant inside every level-0 block
    order = np.random.default_rng(0).permutation(len(db))       # level-0 score is not monotone in the row id
    db = db[order]
    q = np.repeat(np.concatenate([[5.0, 3.0, 2.0, 1.0], np.full(20, 2.0)]).astype(np.float32), 64)
    d = hq.EmbeddingDatabase(db, n=n)
    want_i, want_s = O.progressive_search(q, db, n, 10)
    ids_ref, sc_ref = hq.search_batch(d, q[None, :], 10, tie_rule="reference")
    m = len(want_i)
    assert m >= 2
    assert list(ids_ref[0, :m].cpu().numpy()) == list(want_i), (ids_ref[0].cpu().numpy(), want_i)
    assert (ids_ref[0, m:] == -1).all()
    assert np.abs(sc_ref[0, :m].cpu().numpy() - want_s).max() < 5e-7
    ids_id, _ = hq.search_batch(d, q[None, :], 10, filter_impl="exact")
    assert sorted(ids_id[0, :m].cpu().numpy()) != sorted(want_i)         # the row-id rule keeps other rows of the tie group
    # the rule costs nothing when nothing ties: random rows give identical results under both rules
    rng = np.random.default_rng(1)
    db2 = rng.standard_normal((3000, D)).astype(np.float32)
    qs2 = rng.standard_normal((5, D)).astype(np.float32)
    d2 = hq.EmbeddingDatabase(db2, n=n)
    a_i, a_s = hq.search_batch(d2, qs2, 10, tie_rule="reference")
    b_i, b_s = hq.search_batch(d2, qs2, 10, filter_impl="exact")
    assert torch.equal(a_i, b_i) and torch.equal(a_s, b_s)
