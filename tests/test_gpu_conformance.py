"""Drop-in proof on the GPU box ("unchanged callers must run green with the shim injected").

1. The reference's OWN eleven hot-path test files and the two pipeline test files run, unmodified, in a
   subprocess where `hilbert_quantization_b200.dropin.install()` has swapped the device classes into the reference
   package (tests/conformance/hq_dropin_plugin.py).  The pass count is compared with the all-reference run of the same
   files in the same environment.
2. `QuantizationPipeline` built by constructor injection (core/pipeline.py:37-69) from device components produces the
   same QuantizedModel as the all-reference pipeline, and reconstructs through the device inverse.
The reference package comes from baseline/_ref (baseline/install_reference.sh; git-ignored, shipped by gpurun)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "conformance"))
import run_reference_tests as conf  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not conf.available(), reason="baseline/_ref missing (run baseline/install_reference.sh)")]

# Reference tests the device classes are allowed to fail, each with the reason (documented in DESIGN.md section 8).
EXPECTED_FAILURES: dict = {}


def _dump(name, rep):
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, name), "w") as f:
            json.dump(rep, f, indent=1)


def test_reference_hot_path_tests_pass_on_the_device_classes():
    native = conf.run(conf.HOT_PATH_FILES, native=True)
    device = conf.run(conf.HOT_PATH_FILES, native=False)
    _dump("conformance_hot_path.json", {"native": {k: native[k] for k in ("passed", "failed", "per_file")},
                                         "device": {k: device[k] for k in ("passed", "failed", "per_file")},
                                         "device_tail": device["tail"]})
    assert native["passed"] >= 289, native["tail"]
    unexpected = [f for f in device["failed"] if f["test"] not in EXPECTED_FAILURES]
    assert not unexpected, "\n".join(f"{f['test']}: {f['message']}" for f in unexpected) + "\n" + device["tail"][-3000:]
    assert device["passed"] + len(device["failed"]) == native["passed"] + len(native["failed"])
    assert device["passed"] >= native["passed"] - len(EXPECTED_FAILURES)


def test_unchanged_pipeline_callers_run_on_the_device_classes():
    native = conf.run(conf.CALLER_FILES, native=True)
    device = conf.run(conf.CALLER_FILES, native=False)
    _dump("conformance_callers.json", {"native": {k: native[k] for k in ("passed", "failed", "per_file")},
                                        "device": {k: device[k] for k in ("passed", "failed", "per_file")},
                                        "device_tail": device["tail"]})
    native_failed = {f["test"] for f in native["failed"]}
    new = [f for f in device["failed"] if f["test"] not in native_failed]       # the reference itself fails one of them
    assert not new, "\n".join(f"{f['test']}: {f['message']}" for f in new) + "\n" + device["tail"][-3000:]
    assert device["passed"] >= native["passed"]


@pytest.fixture(scope="module")
def ref_modules():
    sys.path.insert(0, conf.REF)
    try:
        import hilbert_quantization.core.pipeline as pl
        import hilbert_quantization.config as cfg
        yield pl, cfg
    finally:
        sys.path.remove(conf.REF)


@pytest.mark.parametrize("count, streaming", [(1024, True), (4096, True), (3000, True), (4096, False), (16384, True)])
def test_pipeline_with_injected_device_components_equals_the_reference_pipeline(ref_modules, count, streaming):
    """core/pipeline.py:37-69 constructor injection; :129-146 map -> index -> embed -> compress."""
    import hilbert_quantization_b200 as hq
    pl, cfg = ref_modules
    rng = np.random.default_rng(count)
    params = rng.standard_normal(count).astype(np.float32)
    qcfg = cfg.QuantizationConfig(use_streaming_optimization=streaming)
    ours = pl.QuantizationPipeline(dimension_calculator=hq.PowerOf4DimensionCalculator(),
                                   hilbert_mapper=hq.HilbertCurveMapper(),
                                   index_generator=hq.HierarchicalIndexGeneratorImpl(qcfg),
                                   use_streaming_optimization=streaming)
    ref = pl.QuantizationPipeline(use_streaming_optimization=streaming) if streaming else \
        pl.QuantizationPipeline(index_generator=pl.HierarchicalIndexGeneratorImpl(qcfg), use_streaming_optimization=False)
    a = ours.quantize_model(params, "m", compression_quality=0.9)
    b = ref.quantize_model(params, "m", compression_quality=0.9)
    assert a.original_dimensions == b.original_dimensions and a.parameter_count == b.parameter_count
    ia, ib = np.asarray(a.hierarchical_indices), np.asarray(b.hierarchical_indices)
    assert ia.dtype == ib.dtype and ia.shape == ib.shape
    if streaming:
        assert np.array_equal(ia, ib)                       # variant B: bit exact
    else:
        assert np.abs(ia - ib).max() <= 3e-7                # variant A: fp32 means, tree vs pairwise order
    if streaming:
        assert a.compressed_data == b.compressed_data       # same image + same index row -> same JPEG bytes
    ra, rb = ours.reconstruct_parameters(a), ref.reconstruct_parameters(b)
    assert ra.shape == rb.shape == (count,)
    if streaming:
        assert np.array_equal(ra, rb)


def test_integrated_mapping_hook_equals_the_separate_calls():
    """generate_indices_with_integrated_mapping (probed by core/pipeline.py:115-121) == map_to_2d + generate_optimized_indices."""
    import hilbert_quantization_b200 as hq

    class Cfg:
        use_streaming_optimization = True
    rng = np.random.default_rng(5)
    for Config in (Cfg, None):
        gen = hq.HierarchicalIndexGeneratorImpl(Config() if Config else None)
        for n, count in ((8, 64), (32, 1000), (64, 4096)):
            p = rng.standard_normal(count).astype(np.float32)
            img, idx = gen.generate_indices_with_integrated_mapping(p, (n, n), n)
            img2 = hq.HilbertCurveMapper().map_to_2d(p, (n, n))
            idx2 = gen.generate_optimized_indices(img2, n)
            assert img.dtype == img2.dtype and np.array_equal(img, img2)
            assert idx.dtype == idx2.dtype and np.array_equal(idx, idx2)
    with pytest.raises(hq.HilbertQuantizationError, match="requires square dimensions"):
        hq.HierarchicalIndexGeneratorImpl().generate_indices_with_integrated_mapping(np.zeros(8, dtype=np.float32), (2, 4), 4)
