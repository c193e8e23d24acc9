"""pytest plugin (`-p hq_dropin_plugin`): swap the device classes into the importable reference package before
its test modules are collected, so the reference's OWN hot-path tests (tests/test_hilbert_mapper.py, ... SURVEY 8c)
run unmodified against the sm_100a implementation.  See hilbert_quantization_b200/dropin.py."""
import os


def pytest_configure(config):
    import hilbert_quantization_b200.dropin as dropin
    classes = dropin.install()
    if os.environ.get("HQ_CONFORMANCE_VERBOSE"):
        print("hq drop-in installed:", ", ".join(sorted(classes)))
