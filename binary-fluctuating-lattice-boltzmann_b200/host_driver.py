"""Build recipe for the C++ host driver (csrc/host/main_run_job.cpp -> bflbm_run_job), linked against libbflbm.so."""
from __future__ import annotations

import os
import subprocess

from . import _build

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "host", "main_run_job.cpp")
EXE = os.path.join(HERE, "bflbm_run_job")


def build(force: bool = False) -> str:
    _build.build()
    _build.build_sf()
    deps = [SRC] + [os.path.join(HERE, "csrc", "host", h) for h in ("bflbm.hpp", "parameters.hpp", "plotfile.hpp")]
    if not force and os.path.exists(EXE) and all(os.path.getmtime(d) <= os.path.getmtime(EXE) for d in deps):
        return EXE
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O2", "-Wall", "-pthread", SRC, "-o", EXE, "-L" + HERE, "-lbflbm_sf", "-lbflbm", "-Wl,-rpath,$ORIGIN",
                    "-Wl,-rpath-link," + HERE, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    return EXE
