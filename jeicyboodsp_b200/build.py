"""Build jeicyboodsp_b200/libjdsp.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).
The translation units are compiled in parallel, then linked."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libjdsp.so")
SOURCES = ["jdsp_api.cu", "jdsp_stft.cu", "jdsp_conv_mfcc.cu", "jdsp_pitch.cu", "jdsp_mvdr.cu"]
COMMON = ["jdsp_host.hpp", "jdsp_device.cuh", "../../include/jdsp.h"]
DEPS = {"jdsp_api.cu": ["kernels_fft.cuh"], "jdsp_stft.cu": ["kernels_stft.cuh", "kernels_stream.cuh"],
        "jdsp_conv_mfcc.cu": ["kernels_conv_mfcc.cuh", "kernels_fastconv.cuh", "kernels_mfcc.cuh", "kernels_stream.cuh", "kernels_stft.cuh"],
        "jdsp_pitch.cu": ["kernels_pitch.cuh", "kernels_stft.cuh"],
        "jdsp_mvdr.cu": ["kernels_mvdr.cuh", "kernels_stft.cuh"]}
HEADERS = COMMON + sorted({h for v in DEPS.values() for h in v})
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libjdsp.so is CUDA-only (sm_100a) and has no CPU build")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _compile(src: str):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    log = obj + ".log"
    deps = [os.path.join(CSRC, f) for f in [src] + COMMON + DEPS[src]]
    if os.path.exists(obj) and os.path.exists(log) and all(
            not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(obj) for d in deps):
        return src, obj, subprocess.CompletedProcess([], 0, "", open(log).read())   # up to date
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", "-o", obj, os.path.join(CSRC, src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode == 0:
        with open(log, "w") as f:
            f.write(r.stderr)
    elif os.path.exists(obj):
        os.remove(obj)
    return src, obj, r


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(len(SOURCES)) as ex:
        results = list(ex.map(_compile, SOURCES))
    log = os.path.join(CSRC, "_ptxas.log")
    with open(log, "w") as f:
        for src, _, r in results:
            f.write(f"==== {src}\n{r.stderr}")
    for src, _, r in results:
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + r.stderr[-4000:])
    # the arch flag at link time too: without it nvcc adds an empty default-arch (sm_52) device-link stub to the fat binary
    r = subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [o for _, o, _ in results],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    if verbose:
        for _, _, rr in results:
            print(rr.stderr[-1500:])
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
