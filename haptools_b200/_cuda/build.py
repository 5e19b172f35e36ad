"""
Builds haptools_b200/_cuda/libhaptools_cuda.so from haptools_b200/csrc/*.cu with nvcc for
sm_100a (B200).  In-tree on purpose: the .so is git-ignored but travels with the repo
snapshot to the GPU box.

    python -m haptools_b200._cuda.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent.parent
REPO = PKG.parent
CSRC = PKG / "csrc"
INCLUDE = REPO / "include"
LIB_PATH = Path(__file__).resolve().parent / "libhaptools_cuda.so"
SOURCES = ["ht_capi.cu", "ht_hostcopy.cu", "ht_expand_host.cpp", "ht_pack.cu", "ht_pack_anc_tma.cu", "ht_transform.cu", "ht_transform_half.cu", "ht_transform_local.cu", "ht_hapmajor.cu", "ht_synth.cu", "ht_breakpoints.cu", "ht_qc.cu"]
NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "--use_fast_math",
    "-Xcompiler",
    "-fPIC,-O2,-Wall",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found: cannot build libhaptools_cuda.so")
    return nvcc


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    built = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    return any(d.stat().st_mtime > built for d in deps)


def build(force: bool = False, verbose: bool = False, defines: list = None, out: Path = None) -> Path:
    """Compile and link the shared library; returns its path.  `defines` / `out`: an experimental variant
    (e.g. ["HT_STAGE_MIN_CTAS=2"]) written next to the product library and loaded through HT_LIB_PATH."""
    lib_path = Path(out) if out else LIB_PATH
    if not force and not defines and not out and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    objdir = REPO / "build" / (lib_path.stem if out else "")
    objdir.mkdir(parents=True, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = objdir / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in (defines or [])], f"-I{INCLUDE}", f"-I{CSRC}", "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out, file=sys.stderr)
        objs.append(str(obj))
    tmp = lib_path.with_suffix(".so.tmp%d" % os.getpid())
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp), *objs, "--cudart", "static"]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}")
    os.replace(tmp, lib_path)
    return lib_path


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs or None, out=outs[0] if outs else None))
