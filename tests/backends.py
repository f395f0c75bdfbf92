"""Two ways to reach the same C ABI from the parity tests:

* ``gpu``  -- jeicyboodsp_b200/libjdsp.so on a CUDA device (the product; tests marked ``gpu``).
* ``emul`` -- the same sources compiled against tests/emul (CPU execution emulator), small sizes only.
  It exists because the build container has no GPU; it is test infrastructure, never a product path.
"""
from __future__ import annotations

import os
import subprocess

import numpy as np

from jeicyboodsp_b200.binding import Context, Library

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL_LIB = os.path.join(HERE, "emul", "_build", "libjdsp_emul.so")


class EmulBackend:
    name = "emul"

    def __init__(self):
        src_dir = os.path.join(os.path.dirname(HERE), "jeicyboodsp_b200", "csrc")
        deps = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh"))]
        deps += [os.path.join(HERE, "emul", f) for f in ("cuda_emul.h", "cuda_emul.cpp")]
        if not os.path.exists(EMUL_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMUL_LIB) for d in deps):
            subprocess.run(["bash", os.path.join(HERE, "emul", "build_emul.sh")], check=True, stdout=subprocess.DEVNULL)
        self.L = Library(EMUL_LIB)
        self.ctx = Context(self.L)

    def set_order(self, order: int) -> None:
        import ctypes as C
        C.c_int.in_dll(self.L.lib, "_ZN9jdsp_emul7g_orderE").value = order

    def zeros(self, shape, dtype):
        return np.zeros(shape, dtype)

    def to_dev(self, a: np.ndarray):
        return np.ascontiguousarray(a).copy()

    def to_host(self, a) -> np.ndarray:
        return np.asarray(a)

    def sync(self) -> None:
        pass


class GpuBackend:
    name = "gpu"

    def __init__(self):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("GPU backend requested without a CUDA device")
        self.torch = torch
        self.L = Library()
        self.ctx = Context(self.L, 0)

    def set_order(self, order: int) -> None:
        pass

    def zeros(self, shape, dtype):
        t = self.torch
        m = {np.int16: t.int16, np.float32: t.float32, np.uint8: t.uint8, np.complex64: t.complex64,
             np.complex128: t.complex128, np.int32: t.int32, np.float64: t.float64}
        return t.zeros(shape, dtype=m[np.dtype(dtype).type], device="cuda")

    def to_dev(self, a: np.ndarray):
        return self.torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def to_host(self, a) -> np.ndarray:
        self.ctx.sync()
        return a.cpu().numpy()

    def sync(self) -> None:
        self.ctx.sync()
        self.torch.cuda.synchronize()
