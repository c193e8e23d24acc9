#!/usr/bin/env bash
# Install the UNMODIFIED reference (pure Python, /root/reference) into baseline/_ref and put its
# eleven hot-path test files next to it.  baseline/_ref is git-ignored (never in
# history) but not gpurun-ignored, so it travels to the GPU box, where /root/reference does
# not exist.  Used by: tests/conformance (the reference's own tests against the device classes),
# tests/test_gpu_dropin.py (QuantizationPipeline with injected device components) and
# `bench.py --impl reference` (cpu_baseline.kind = "reference").
#
# Outcome recorded in DESIGN.md: plain `pip install --target` fails on dependency resolution
# (numpy is not in the offline wheelhouse: it is already installed in the image), `--no-deps`
# from a writable copy succeeds (setuptools writes an egg-info into the source tree).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${HQ_REFERENCE_PATH:-/root/reference}"
[ -d "$REF/hilbert_quantization" ] || { echo "no reference checkout at $REF" >&2; exit 1; }
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF" "$TMP/ref"
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP/ref" >/dev/null
mkdir -p "$HERE/_ref/ref_tests"
for t in test_hilbert_mapper test_rag_hilbert_mapper test_inverse_mapping_requirements test_index_generator \
         test_streaming_index_generator test_hierarchical_index_generator test_hierarchical_index_comparison \
         test_progressive_filtering test_similarity_calculation test_search_engine test_dimension_calculator \
         test_quantization_pipeline test_reconstruction_pipeline; do
    cp "$REF/tests/$t.py" "$HERE/_ref/ref_tests/"
done
echo "reference installed into $HERE/_ref ($(ls "$HERE/_ref/ref_tests" | wc -l) test files)"
