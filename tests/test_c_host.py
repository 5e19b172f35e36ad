"""The drop-in boundary from a host that is not Python: tests/c_host/abi_roundtrip.c is plain C99
and sees nothing but include/haptools_cuda.h and the CUDA runtime's C API.  CPU suite: the header
is valid C99 and C++, the program compiles and links against libhaptools_cuda.so.  GPU suite: it
runs the pack + transform path in every pack mode and compares the result bytes with the
definition of the transform evaluated by its own loops."""
import shutil
import subprocess
from pathlib import Path

import pytest

from haptools_b200 import _cuda
from haptools_b200._cuda import build as cuda_build

REPO = Path(__file__).resolve().parent.parent
HEADER = REPO / "include" / "haptools_cuda.h"
SOURCE = REPO / "tests" / "c_host" / "abi_roundtrip.c"
CUDA_HOME = Path("/usr/local/cuda")


def _compile(out: Path) -> Path:
    cuda_build.build()  # no-op when up to date
    libdir = _cuda.LIB_PATH.parent
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O2", f"-I{REPO / 'include'}", f"-I{CUDA_HOME / 'include'}",
           str(SOURCE), "-o", str(out), f"-L{libdir}", "-lhaptools_cuda", f"-L{CUDA_HOME / 'lib64'}", "-lcudart",
           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{CUDA_HOME / 'lib64'}"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    return out


needs_gcc = pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not found")


@needs_gcc
@pytest.mark.parametrize("compiler,lang", [("gcc", ["-std=c99", "-x", "c"]), ("g++", ["-std=c++11", "-x", "c++"])])
def test_header_is_plain_c_and_cxx(compiler, lang):
    res = subprocess.run([compiler, *lang, "-Wall", "-Wextra", "-Werror", "-fsyntax-only", str(HEADER)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout


@needs_gcc
def test_c_host_compiles_and_links(tmp_path):
    exe = _compile(tmp_path / "abi_roundtrip")
    assert exe.exists()


@needs_gcc
@pytest.mark.gpu
@pytest.mark.parametrize("n,p,H,seed", [(2504, 96, 300, 1), (64, 8, 5, 2), (1025, 33, 17, 3), (3000, 200, 129, 4), (1, 1, 1, 5)])
def test_c_host_roundtrip(tmp_path, n, p, H, seed):
    exe = _compile(tmp_path / "abi_roundtrip")
    res = subprocess.run([str(exe), str(n), str(p), str(H), str(seed)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=120)
    assert res.returncode == 0, res.stdout
    assert "all checks passed" in res.stdout and "DIFFER" not in res.stdout, res.stdout
