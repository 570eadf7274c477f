"""Import the *unmodified* reference sampler from /root/reference.  TEST INFRASTRUCTURE.

Only usable in the build container (the GPU box has no /root/reference); used by
oracle/gen_golden.py to pin the restatement and to generate tests/golden/*.pt.
`matplotlib`, `normflows` and `nflows` are absent here, so they are stubbed in
sys.modules before `import fab` (recipe from SURVEY Appendix C).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FAB_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fab"))


def load_reference():
    """Returns the imported `fab` package of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    plt = sys.modules["matplotlib.pyplot"]
    plt.Figure = plt.Axes = object
    if "normflows" not in sys.modules:
        nf = types.ModuleType("normflows")
        nf.NormalizingFlow = object
        sys.modules["normflows"] = nf
    if "nflows" not in sys.modules:
        nfl = types.ModuleType("nflows")
        nfl.flows = types.SimpleNamespace(Flow=object)
        sys.modules["nflows"] = nfl
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import fab  # noqa: F401
    return fab
