"""The numpy prototype of the cluster-moment far field (tools/proto/m2l_lorentz.py) keeps its
identities: P2M recurrence, suffix-sum moment-to-local translation, truncation below 1e-9 against the direct sum."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_m2l_prototype_matches_direct_sum():
    spec = importlib.util.spec_from_file_location("m2l_lorentz", os.path.join(ROOT, "tools", "proto", "m2l_lorentz.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()      # asserts max relative error < 1e-9 for every (theta, p) it tries
