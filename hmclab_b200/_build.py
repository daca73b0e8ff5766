"""In-tree build of the CUDA engine: ``hmclab_b200/lib/libhmcb.so`` (sm_100a only).

``python -m hmclab_b200._build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the kernel families live in separate translation units (and the widest
template families are split further by ``-D`` parameters) so they compile in parallel.
The shared library is git-ignored but travels with the working tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libhmcb.so")
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    # elementwise arithmetic is written with explicit __dmul_rn/__dadd_rn; keep the few
    # plain expressions (RNG transforms) uncontracted as well
    "-fmad=false",
]

# (source, object suffix, extra defines)
UNITS = (
    [("hmcb.cu", "", []), ("launch_fused.cu", "", []), ("launch_srcloc.cu", "", []),
     ("launch_staged.cu", "", []), ("launch_spmm.cu", "", []), ("launch_fused_dense.cu", "", []),
     ("debug_peak.cu", "", []), ("launch_ozaki.cu", "", [])]
    + [("launch_fused_ppt.cu", f"_{p}", [f"-DHMCB_PPT={p}"]) for p in (1, 2, 4)]
    + [("launch_srcloc_lpe.cu", f"_{l}_{n}", [f"-DHMCB_LPE={l}", f"-DHMCB_NP={n}"])
       for l in (1, 2, 4) for n in (3, 4)]
)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the B200 engine cannot be built")


def _source_digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join(INCLUDE, "hmcb.h")]
    for name in files:
        path = name if os.path.isabs(name) else os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _unit_digest(src: str, defines) -> str:
    """Hash of a translation unit: its source, the closure of its local #include files, flags."""
    import re
    seen, todo = set(), [os.path.join(CSRC, src)]
    while todo:
        path = os.path.normpath(todo.pop())
        if path in seen or not os.path.isfile(path):
            continue
        seen.add(path)
        with open(path) as f:
            for inc in re.findall(r'^\s*#include\s+"([^"]+)"', f.read(), flags=re.M):
                todo.append(os.path.join(os.path.dirname(path), inc))
    h = hashlib.sha256()
    for path in sorted(seen):
        h.update(path.encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(list(NVCC_FLAGS) + list(defines)).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = LIB_PATH + ".digest"
    if not (os.path.exists(LIB_PATH) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == _source_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if sources changed) and return the path of libhmcb.so."""
    if not force and is_current():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)

    def compile_unit(unit):
        src, suffix, defines = unit
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + suffix + ".o")
        digest, stamp = _unit_digest(src, defines), obj + ".digest"
        if not force and os.path.exists(obj) and os.path.exists(stamp):
            with open(stamp) as f:
                if f.read().strip() == digest:
                    return obj
        cmd = [nvcc, *NVCC_FLAGS, *defines, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}{suffix}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        with open(stamp, "w") as f:
            f.write(digest)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as pool:
        objects = list(pool.map(compile_unit, UNITS))
    link = [nvcc, "-shared", "-o", LIB_PATH, *objects, "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(LIB_PATH + ".digest", "w") as f:
        f.write(_source_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
