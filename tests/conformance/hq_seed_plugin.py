"""pytest plugin (`-p hq_seed_plugin`), used for the native AND the device run of the reference's tests: NumPy's and
Python's global random generators are seeded from the test's node id before every test.  Several reference tests draw
unseeded `np.random.randn` inputs and push them through lossy JPEG with fixed error thresholds
(e.g. tests/test_reconstruction_pipeline.py::test_index_validation); unseeded, the same test passes or fails from run
to run, and the native / device comparison would compare two different inputs.  The test files stay unmodified."""
import random
import zlib

import numpy as np
import pytest


@pytest.fixture(autouse=True)
def _hq_seed_from_node_id(request):
    seed = zlib.crc32(request.node.nodeid.encode()) & 0x7FFFFFFF
    np.random.seed(seed)
    random.seed(seed)
    yield
