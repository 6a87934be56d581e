"""TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference `_cmatrices` C extension.

Compiles /root/reference/radiomics/src/{_cmatrices.c,cmatrices.c} where they lie (nothing is
copied into this repository) into `oracle/_ref/_cmatrices<EXT_SUFFIX>`, with a plain gcc
command.  `oracle/_ref/` is git-ignored but travels to the GPU box,
where it serves (a) as the matrix-level checker of the CUDA path and (b) as the
`cpu_baseline` / `--impl reference` arm of bench.py.

Only tests/, __graft_entry__.smoke()/build() and bench.py's cpu legs may use this.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/radiomics/src"
OUT_DIR = os.path.join(HERE, "_ref")


MODULES = {"_cmatrices": ["_cmatrices.c", "cmatrices.c"], "_cshape": ["_cshape.c", "cshape.c"]}


def ref_so_path(name: str = "_cmatrices") -> str:
    return os.path.join(OUT_DIR, name + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force: bool = False, name: str = "_cmatrices") -> str | None:
    """Return the path of the built module, or None when the reference sources are absent
    (GPU box) and no prebuilt file exists."""
    out = ref_so_path(name)
    srcs = [os.path.join(REF_SRC, f) for f in MODULES[name]]
    if not all(os.path.exists(s) for s in srcs):
        return out if os.path.exists(out) else None
    if os.path.exists(out) and not force:
        if os.path.getmtime(out) >= max(os.path.getmtime(s) for s in srcs):
            return out
    import numpy

    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [
        "gcc", "-O2", "-shared", "-fPIC",
        "-I" + sysconfig.get_paths()["include"],
        "-I" + numpy.get_include(),
        "-I" + REF_SRC,
        *srcs, "-o", out,
    ]
    subprocess.check_call(cmd)
    return out


def load(name: str = "_cmatrices"):
    """Import the compiled reference module as a standalone module object."""
    import importlib.util

    path = build(name=name)
    if path is None:
        raise ImportError(f"reference {name} not built and /root/reference absent")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p)
