"""Build recipe for oracle/_ref: the UNMODIFIED reference C extension.

TEST INFRASTRUCTURE ONLY.  Compiles the reference's own three C files
(setup.py:29-33 of the reference: _deform_grid.c, deform.c, from_nd_image.c)
where they lie under /root/reference with plain gcc -O2 (the flags setuptools
uses minus the distro hardening), and writes ONLY the resulting extension
module into oracle/_ref/.  No reference source is copied into this repo.

oracle/_ref/ is git-ignored but NOT gpurun-ignored, so the built .so travels
to the GPU box where /root/reference does not exist.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("EDF_REFERENCE_SRC", "/root/reference/elasticdeform")
OUT_DIR = os.path.join(HERE, "_ref")
SOURCES = ["_deform_grid.c", "deform.c", "from_nd_image.c"]


def ref_so_path():
    return os.path.join(OUT_DIR, "_deform_grid" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False, verbose=False):
    """Compile the reference extension. Returns the .so path, or None when the
    reference sources are not present (e.g. on the GPU box: use the prebuilt file)."""
    so = ref_so_path()
    srcs = [os.path.join(REF_SRC, s) for s in SOURCES]
    if not all(os.path.exists(s) for s in srcs):
        return so if os.path.exists(so) else None
    if os.path.exists(so) and not force:
        if all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
            return so
    import numpy
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fwrapv", "-w",
           "-I", REF_SRC, "-I", numpy.get_include(),
           "-I", sysconfig.get_paths()["include"],
           *srcs, "-o", so, "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref:", p)
