"""Build libacvm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m acvm_b200.build [--force]

Objects are cached under acvm_b200/csrc/_obj (git-ignored); the .so lands next to the package so
it travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libacvm_b200.so")
SOURCES = ["vm_kernel_full.cu", "vm_kernel_full_b.cu", "vm_kernel.cu", "runtime.cu", "acir.cpp", "plan.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"]


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".hpp", ".h", ".cu", ".cpp")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, force):
    obj = os.path.join(OBJ, src + ".o")
    srcp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= _deps_mtime():
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-Xptxas", "-v", "-c", srcp, "-o", obj]
    if src.endswith(".cpp"):
        cmd = [NVCC] + FLAGS + ["-x", "cu", "-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=6) as ex:
        res = list(ex.map(lambda s: _compile(s, force), SOURCES))
    objs = [o for o, _ in res]
    log = "\n".join(l for _, l in res if l)
    if verbose and log:
        print(log)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
            f.write(log)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        # --no-undefined: a symbol missing from the objects must fail HERE, not at the first call on the GPU box
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lz", "-lpthread", "-ldl", "-lrt", "-Xlinker", "--no-undefined",
                                                      "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
