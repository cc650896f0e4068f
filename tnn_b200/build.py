"""In-tree native build for tnn_b200.

Compiles every ``csrc/kernels/*.cu`` with nvcc for sm_100a and every
``csrc/**/*.cpp`` with g++, then links ``tnn_b200/_C.so`` (a CPython extension
bound with pybind11 through ``torch/extension.h``).  Object files are cached by
content hash under ``build/`` so rebuilding after a one-file edit takes seconds.
nvcc cross-compiles, so this works on a box without a GPU; the resulting ``.so``
travels with the tree to the GPU box.

Usage:  python -m tnn_b200.build [--force] [--jobs N]
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
BUILD = ROOT / "_build"   # travels with the tree (gpurun) so the GPU box does not rebuild
TARGET = ROOT / "_C.so"

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _cuda_home() -> str:
    for c in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if c and os.path.exists(os.path.join(c, "bin", "nvcc")):
            return c
    nv = shutil.which("nvcc")
    if nv:
        return str(Path(nv).resolve().parent.parent)
    raise RuntimeError("nvcc not found")


def _torch_paths():
    import torch
    from torch.utils import cpp_extension as ce

    inc = ce.include_paths()
    lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    return inc, lib, int(torch._C._GLIBCXX_USE_CXX11_ABI)


def _sources():
    cu = sorted((CSRC / "kernels").glob("*.cu"))
    cpp = sorted(CSRC.rglob("*.cpp"))
    return cu, cpp


def _headers_digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.rglob("*.cuh")) + list(CSRC.rglob("*.hpp")) + list(CSRC.rglob("*.h"))):
        h.update(p.read_bytes())
    return h.hexdigest()


def _flag_sig(cmd: list[str]) -> str:
    """Flags that affect code generation, minus absolute paths (the tree is relocated on the GPU box)."""
    return " ".join(c for c in cmd[1:] if not c.startswith("-I") and not c.startswith("/"))


def _compile_one(src: Path, cmd: list[str], obj: Path, stamp: str) -> tuple[Path, bool, str]:
    stamp_file = obj.with_suffix(".stamp")
    if obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return obj, False, ""
    r = subprocess.run(cmd + ["-c", str(src), "-o", str(obj)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"compile failed: {' '.join(cmd)} {src}\n{r.stdout}\n{r.stderr}")
    stamp_file.write_text(stamp)
    return obj, True, r.stderr


def build(force: bool = False, jobs: int | None = None, verbose: bool = True) -> Path:
    cuda = _cuda_home()
    inc, torch_lib, cxx11 = _torch_paths()
    py_inc = sysconfig.get_paths()["include"]
    BUILD.mkdir(parents=True, exist_ok=True)
    cu, cpp = _sources()
    hdr = _headers_digest()
    nvcc = [
        os.path.join(cuda, "bin", "nvcc"), *NVCC_ARCH, "-O3", "-std=c++17", "-lineinfo",
        "--use_fast_math", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
        f"-I{CSRC}", f"-I{CSRC / 'kernels'}",
    ]
    gxx = [
        "g++", "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-Wno-deprecated-declarations",
        f"-D_GLIBCXX_USE_CXX11_ABI={cxx11}", "-DTORCH_EXTENSION_NAME=_C",
        "-DTORCH_API_INCLUDE_EXTENSION_H", f"-I{CSRC}", f"-I{CSRC / 'kernels'}",
        f"-I{cuda}/include", f"-I{py_inc}", *[f"-I{p}" for p in inc], "-pthread",
    ]
    tasks = []
    for s in cu:
        stamp = hashlib.sha256(s.read_bytes() + hdr.encode() + _flag_sig(nvcc).encode()).hexdigest()
        tasks.append((s, nvcc, BUILD / (s.stem + ".cu.o"), "" if force else stamp, stamp))
    for s in cpp:
        stamp = hashlib.sha256(s.read_bytes() + hdr.encode() + _flag_sig(gxx).encode()).hexdigest()
        tasks.append((s, gxx, BUILD / (s.stem + ".cpp.o"), "" if force else stamp, stamp))
    objs, rebuilt = [], False
    with cf.ThreadPoolExecutor(max_workers=jobs or min(8, os.cpu_count() or 4)) as ex:
        futs = []
        for s, cmd, obj, chk, stamp in tasks:
            if force and obj.with_suffix(".stamp").exists():
                obj.with_suffix(".stamp").unlink()
            futs.append(ex.submit(_compile_one, s, cmd, obj, stamp))
        for f in futs:
            obj, did, _ = f.result()
            objs.append(obj)
            rebuilt |= did
            if did and verbose:
                print(f"[tnn_b200.build] compiled {obj.name}", flush=True)
    if rebuilt or not TARGET.exists():
        link = [
            "g++", "-shared", "-o", str(TARGET), *map(str, objs),
            f"-L{torch_lib}", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10", "-ltorch_cuda",
            "-lc10_cuda", f"-L{cuda}/lib64", "-lcudart", f"-Wl,-rpath,{torch_lib}",
            f"-Wl,-rpath,{cuda}/lib64", "-pthread",
        ]
        r = subprocess.run(link, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[tnn_b200.build] linked {TARGET}", flush=True)
    return TARGET


# ----------------------------------------------------------------------------------------------- C++ host
CPP = ROOT.parent / "cpp"


def build_host(force: bool = False, verbose: bool = True) -> dict:
    """libtnn_b200.so (the C++ framework: cpp/include/tnn + cpp/src, linked with the SAME sm_100a kernel objects as
    _C.so but with NO torch and NO Python) and its executables: cpp/bin/trainer, cpp/bin/bench_cpp, cpp/bin/tnn_tests.
    ``ldd cpp/bin/trainer`` shows libtnn_b200, libcudart, libstdc++ -- nothing else."""
    cuda = _cuda_home()
    build(force=False, verbose=verbose)                       # kernel objects (tnn_b200/_build/*.cu.o)
    obj_dir, lib_dir, bin_dir = CPP / "_build", CPP / "lib", CPP / "bin"
    for d in (obj_dir, lib_dir, bin_dir):
        d.mkdir(parents=True, exist_ok=True)
    hdr = hashlib.sha256()
    for h in sorted(list((CPP / "include").rglob("*.hpp")) + list((CPP / "examples").glob("*.hpp")) + list(CSRC.rglob("*.hpp")) +
                    list(CSRC.rglob("*.cuh"))):
        hdr.update(h.read_bytes())
    gxx = ["g++", "-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-function", f"-I{CPP / 'include'}", f"-I{CSRC}",
           f"-I{cuda}/include", "-pthread"]
    tasks = []
    for s in sorted((CPP / "src").glob("*.cpp")) + sorted((CPP / "examples").glob("*.cpp")) + sorted((CPP / "tests").glob("*.cpp")):
        stamp = hashlib.sha256(s.read_bytes() + hdr.hexdigest().encode() + _flag_sig(gxx).encode()).hexdigest()
        tasks.append((s, obj_dir / (s.stem + ".o"), stamp))
    rebuilt = False
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        futs = []
        for s, obj, stamp in tasks:
            if force and obj.with_suffix(".stamp").exists():
                obj.with_suffix(".stamp").unlink()
            futs.append(ex.submit(_compile_one, s, gxx, obj, stamp))
        for f in futs:
            obj, did, _ = f.result()
            rebuilt |= did
            if did and verbose:
                print(f"[tnn_b200.build] compiled cpp/{obj.name}", flush=True)
    lib = lib_dir / "libtnn_b200.so"
    lib_objs = [obj_dir / (s.stem + ".o") for s in sorted((CPP / "src").glob("*.cpp"))]
    cu_objs = sorted(BUILD.glob("*.cu.o"))
    # link decisions look at the objects' mtimes, not only at "compiled in this call": an interrupted build leaves fresh
    # objects (with their stamps) next to a stale library / executable
    newest_in = max(o.stat().st_mtime for o in cu_objs + lib_objs)
    if rebuilt or not lib.exists() or lib.stat().st_mtime < newest_in:
        link = ["g++", "-shared", "-o", str(lib), *map(str, lib_objs), *map(str, cu_objs), f"-L{cuda}/lib64", "-lcudart",
                f"-Wl,-rpath,{cuda}/lib64", "-pthread", "-ldl"]
        r = subprocess.run(link, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"libtnn_b200 link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[tnn_b200.build] linked {lib}", flush=True)
        rebuilt = True
    # Python binding of the C++ host (pybind11 only -- no torch): tnn_b200/_host.so
    import pybind11
    host_so = ROOT / "_host.so"
    bsrc = CPP / "python" / "host_bindings.cpp"
    bobj = obj_dir / "host_bindings.o"
    bcmd = gxx + [f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}", "-fvisibility=hidden"]
    bstamp = hashlib.sha256(bsrc.read_bytes() + hdr.hexdigest().encode() + _flag_sig(bcmd).encode()).hexdigest()
    _, bdid, _ = _compile_one(bsrc, bcmd, bobj, bstamp)
    if bdid or rebuilt or not host_so.exists() or host_so.stat().st_mtime < max(bobj.stat().st_mtime, lib.stat().st_mtime):
        link = ["g++", "-shared", "-o", str(host_so), str(bobj), f"-L{lib_dir}", "-ltnn_b200", f"-L{cuda}/lib64", "-lcudart",
                "-Wl,-rpath,$ORIGIN/../cpp/lib", f"-Wl,-rpath,{cuda}/lib64", "-pthread"]
        r = subprocess.run(link, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"_host.so link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[tnn_b200.build] linked {host_so}", flush=True)
    exes = {"_host": host_so}
    for s in sorted((CPP / "examples").glob("*.cpp")) + sorted((CPP / "tests").glob("*.cpp")):
        exe = bin_dir / s.stem
        if rebuilt or not exe.exists() or exe.stat().st_mtime < max((obj_dir / (s.stem + ".o")).stat().st_mtime, lib.stat().st_mtime):
            link = ["g++", "-o", str(exe), str(obj_dir / (s.stem + ".o")), f"-L{lib_dir}", "-ltnn_b200", f"-L{cuda}/lib64", "-lcudart",
                    "-Wl,-rpath,$ORIGIN/../lib", f"-Wl,-rpath,{cuda}/lib64", "-pthread"]
            r = subprocess.run(link, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"link of {exe.name} failed:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(f"[tnn_b200.build] linked {exe}", flush=True)
        exes[s.stem] = exe
    return {"lib": lib, **exes}


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    if "--no-host" not in sys.argv:
        build_host(force="--force" in sys.argv)
