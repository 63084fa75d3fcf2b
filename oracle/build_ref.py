"""Stage the UNMODIFIED reference package for the GPU box  --  TEST / BENCH INFRASTRUCTURE ONLY.

The reference is pure Python (nothing to compile).  ``pip install --no-index --no-build-isolation --no-deps --target
baseline/_ref /root/reference`` fails in this image because its ``setup.py`` requires ``setuptools_scm``, which is not
in the offline wheelhouse (recorded in DESIGN.md), so this script does by hand what that install would have done: it
archives the package directory ``topo_descriptors/`` (the .py files and ``config/``), byte for byte, into
``oracle/_ref/topo_descriptors_ref.zip`` (importable as is through zipimport).  ``oracle/_ref/`` is git-ignored (no reference source enters the history) but not gpurun-ignored, so
it travels to the GPU box, where ``/root/reference`` does not exist; ``bench.py``'s CPU legs import it from there
(``oracle/ref_runner.py``) together with the four stub modules of ``oracle/_stubs``.

    python oracle/build_ref.py            # run by __graft_entry__.build() when /root/reference is present
"""

import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/topo_descriptors"
DST = os.path.join(HERE, "_ref", "topo_descriptors_ref.zip")


def _members():
    out = []
    for base, dirs, files in os.walk(SRC):
        dirs[:] = sorted(d for d in dirs if d != "__pycache__")
        for f in sorted(files):
            if not f.endswith(".pyc"):
                full = os.path.join(base, f)
                out.append((full, os.path.join("topo_descriptors", os.path.relpath(full, SRC))))
    return out


def build(verbose=False):
    """Archive the package if the reference tree is here; returns the staged archive path or None."""
    if not os.path.isdir(SRC):
        return DST if os.path.isfile(DST) else None
    members = _members()
    if os.path.isfile(DST):
        try:
            with zipfile.ZipFile(DST) as z:
                if sorted(z.namelist()) == sorted(n for _, n in members) and all(
                        z.read(n) == open(f, "rb").read() for f, n in members):
                    return DST
        except zipfile.BadZipFile:
            pass
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    with zipfile.ZipFile(DST, "w", zipfile.ZIP_DEFLATED) as z:
        for full, name in members:
            z.write(full, name)
    if verbose:
        print(f"staged {SRC} -> {DST}", file=sys.stderr)
    return DST


if __name__ == "__main__":
    print(build(verbose=True))
