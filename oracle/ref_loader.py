"""Import the *unmodified* reference (lollcat/fab-torch).  TEST / BASELINE INFRASTRUCTURE.

Two places hold it:
  /root/reference      the read-only source tree of the build container (used by
                       oracle/gen_golden.py to pin the restatement and write tests/golden/*.pt);
  baseline/_ref        `pip install --no-deps --target baseline/_ref` of that tree (recipe:
                       __graft_entry__.build(); git-ignored, travels to the GPU box) -- used by
                       `bench.py --impl reference` and by the GPU drop-in tests that run the
                       reference's own fab/core.py on top of the B200 plugin classes.
`matplotlib`, `normflows` and `nflows` are not installable here, so they are stubbed in sys.modules
before `import fab` (recipe from SURVEY Appendix C); nothing on the sampler path touches them.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED_ROOT = os.path.join(_REPO, "baseline", "_ref")


def _pick_root() -> str:
    env = os.environ.get("FAB_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/fab"):
        return "/root/reference"
    return INSTALLED_ROOT


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fab"))


def installed_reference_available() -> bool:
    return os.path.isdir(os.path.join(INSTALLED_ROOT, "fab"))


def install_reference(source: str = "/root/reference") -> bool:
    """pip-install the unmodified reference into baseline/_ref (the source tree is read-only, so
    the wheel is built from a copy under /tmp; --no-deps: normflows & co. are not in the wheelhouse).
    Returns True if baseline/_ref/fab exists afterwards."""
    import shutil
    import subprocess
    import tempfile
    if installed_reference_available():
        return True
    if not os.path.isdir(os.path.join(source, "fab")):
        return False
    tmp = tempfile.mkdtemp(prefix="fab_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(source, src)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", INSTALLED_ROOT, src]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return installed_reference_available()


def load_reference(installed: bool = False):
    """Returns the imported `fab` package of the reference (`installed`: from baseline/_ref)."""
    root = INSTALLED_ROOT if installed else REFERENCE_ROOT
    if not os.path.isdir(os.path.join(root, "fab")):
        raise RuntimeError(f"reference tree not found at {root}")
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
    if root not in sys.path:
        sys.path.insert(0, root)
    import fab  # noqa: F401
    return fab
