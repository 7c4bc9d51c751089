"""Recipe for oracle/_ref/: the UNMODIFIED reference modules of the hot path, copied where they lie under
/root/reference into a git-ignored directory that travels to the GPU box (the reference is pure Python: there is
nothing to compile).  Test infrastructure only -- imported by tests/, smoke() and bench.py's reference arm, never by
the product.

    python oracle/build_ref.py        # no-op when /root/reference is absent (GPU box: uses the prebuilt copy)
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/repellency"
OUT = os.path.join(HERE, "_ref", "repellency")
FILES = ("repellency_methods_fast.py", "repellency_methods_fast_sdv3.py", "repellency_methods_threshold.py",
         os.path.join("utils", "lshash_torch.py"))     # the three modules import it at the top (dead `lsh` method)


def build_ref():
    """Returns the directory holding the reference modules, or None when neither source nor copy exists."""
    if os.path.isdir(REF):
        os.makedirs(OUT, exist_ok=True)
        os.makedirs(os.path.join(OUT, "utils"), exist_ok=True)
        for f in FILES:
            shutil.copyfile(os.path.join(REF, f), os.path.join(OUT, f))
        open(os.path.join(OUT, "__init__.py"), "w").close()
        open(os.path.join(OUT, "utils", "__init__.py"), "w").close()
    return os.path.dirname(OUT) if all(os.path.exists(os.path.join(OUT, f)) for f in FILES) else None


def load(kind="fast"):
    """Import repellency_methods_<kind> of the reference from oracle/_ref (None if it is not there)."""
    import importlib
    import sys
    root = build_ref()
    if root is None:
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    # the reference package is called `repellency` like ours: import it under its own top-level path only
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "repellency" or k.startswith("repellency.")}
    try:
        return importlib.import_module(f"repellency.repellency_methods_{kind}")
    finally:
        for k in [k for k in sys.modules if k == "repellency" or k.startswith("repellency.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)


if __name__ == "__main__":
    print(build_ref())
