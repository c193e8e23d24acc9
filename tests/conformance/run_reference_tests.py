"""Run the reference's own hot-path test files against the device classes.

    python tests/conformance/run_reference_tests.py [--native] [--out report.json] [pytest args / test files]

`--native` runs the same files against the unmodified reference (the CPU figure the device run is compared with).
The reference package and its test files come from baseline/_ref (baseline/install_reference.sh)."""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile
import xml.etree.ElementTree as ET

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "baseline", "_ref")
HOT_PATH_FILES = ["test_hilbert_mapper.py", "test_rag_hilbert_mapper.py", "test_inverse_mapping_requirements.py",
                  "test_index_generator.py", "test_streaming_index_generator.py", "test_hierarchical_index_generator.py",
                  "test_hierarchical_index_comparison.py", "test_progressive_filtering.py", "test_similarity_calculation.py",
                  "test_search_engine.py", "test_dimension_calculator.py"]
CALLER_FILES = ["test_quantization_pipeline.py", "test_reconstruction_pipeline.py"]     # unchanged callers


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "hilbert_quantization")) and os.path.isdir(os.path.join(REF, "ref_tests"))


def run(files=None, native: bool = False, extra=(), timeout: int = 1500) -> dict:
    files = list(files or HOT_PATH_FILES + CALLER_FILES)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([REF, ROOT, HERE] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    with tempfile.TemporaryDirectory() as tmp:
        xml = os.path.join(tmp, "junit.xml")
        cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", REF, f"--junitxml={xml}",
               "-o", "junit_family=xunit1", "--tb=short", "-p", "hq_seed_plugin", *extra]
        if not native:
            cmd += ["-p", "hq_dropin_plugin"]
        cmd += [os.path.join(REF, "ref_tests", f) for f in files]
        p = subprocess.run(cmd, env=env, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
        report = {"native": native, "returncode": p.returncode, "passed": 0, "failed": [], "errors": [], "skipped": 0,
                  "per_file": {}, "tail": p.stdout[-6000:]}
        if os.path.exists(xml):
            for case in ET.parse(xml).getroot().iter("testcase"):
                name = f"{case.get('classname', '')}::{case.get('name')}"
                fname = case.get("file") or case.get("classname", "").split(".")[0]
                per = report["per_file"].setdefault(os.path.basename(fname), {"passed": 0, "failed": 0})
                bad = case.find("failure") if case.find("failure") is not None else case.find("error")
                if bad is not None:
                    report["failed"].append({"test": name, "message": (bad.get("message") or "")[:300]})
                    per["failed"] += 1
                elif case.find("skipped") is not None:
                    report["skipped"] += 1
                else:
                    report["passed"] += 1
                    per["passed"] += 1
        return report


if __name__ == "__main__":
    args = sys.argv[1:]
    native = "--native" in args
    out = None
    if "--out" in args:
        out = args[args.index("--out") + 1]
        del args[args.index("--out"): args.index("--out") + 2]
    args = [a for a in args if a != "--native"]
    files = [a for a in args if a.endswith(".py")] or None
    extra = [a for a in args if not a.endswith(".py")]
    if not available():
        sys.exit("baseline/_ref is missing: run baseline/install_reference.sh where /root/reference exists")
    rep = run(files, native=native, extra=extra)
    print(rep["tail"])
    print(json.dumps({k: rep[k] for k in ("native", "passed", "skipped", "per_file")}, indent=1))
    for f in rep["failed"]:
        print("FAILED", f["test"], "--", f["message"])
    if out:
        with open(out, "w") as fh:
            json.dump(rep, fh, indent=1)
