"""In-tree build of libyoho_b200.so with nvcc for sm_100a (no torch cpp_extension, no JIT cache).

    python -m yoho_b200.build [--force] [-v]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys
import shutil
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libyoho_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# estimator.cu, metrics.cu: FP64 operations must round exactly as written (bit-parity with oracle/estimator_oracle.c)
SOURCES = {
    "abi.cu": [],
    "gconv_simt.cu": [],
    "gconv_tc.cu": [],
    "part1.cu": [],
    "part2.cu": [],
    "match.cu": [],
    "match_tc.cu": [],
    "lift.cu": [],
    "fourier_tc.cu": [],
    "pair.cu": [],
    "train.cu": [],
    "estimator.cu": ["-fmad=false"],
    "metrics.cu": ["-fmad=false"],
}


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "yoho_b200.h"))
    jobs = []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            jobs.append((s, o, extra))

    def run(job):
        s, o, extra = job
        cmd = [_nvcc()] + ARCH + COMMON + extra + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        with open(o + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
        return o

    if jobs:
        with ThreadPoolExecutor(max_workers=4) as ex:
            list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart_static", "-lrt", "-lpthread", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
