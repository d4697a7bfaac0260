"""Every public top-level name of the reference exists in this package (or is a documented absence) — needs the reference
checkout, skipped where it is not mounted."""
import importlib.util
import os

import pytest

REFERENCE = os.environ.get("NXD_REFERENCE_DIR", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src", "neuronx_distributed")), reason="reference checkout not available")
def test_no_unexplained_missing_public_names():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("api_diff", os.path.join(root, "tools", "api_diff.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    missing = mod.diff(REFERENCE, os.path.join(root, "neuronx_distributed_b200"))
    assert not missing, missing
