"""Wall time of the tensor elongation on a full 240x240x155 volume (8.9 M tensors): CUDA kernel through the
host-buffer C ABI (copies included) vs the reference's float32 torch.linalg.eigh on the host cores."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tumorgrowthtoolkit_b200.FK_DTI import tools

shape = (240, 240, 155)
rng = np.random.default_rng(0)
n = int(np.prod(shape))
q, _ = np.linalg.qr(rng.normal(size=(n, 3, 3)))
lam = np.sort(rng.uniform(0.2, 1.0, (n, 3)), axis=-1) * np.array([1.0, 1.5, 2.4]) * 1e-3
t = np.einsum("nij,nj,nkj->nik", q, lam, q)
t = (0.5 * (t + np.swapaxes(t, -1, -2))).reshape(shape + (3, 3))
tools.elongate_tensor_along_main_axis_cuda(t[:2], 2.0)           # context + library load
t0 = time.perf_counter(); a = tools.elongate_tensor_along_main_axis_cuda(t, 2.0); t1 = time.perf_counter()
nsub = n // 8
b = tools.elongate_tensor_along_main_axis_torch(t.reshape(-1, 3, 3)[:nsub], 2.0); t2 = time.perf_counter()
err = float(np.max(np.abs(a.reshape(-1, 3, 3)[:nsub] - b)))
import torch
print(f"cuda (host buffers, {n} tensors): {t1 - t0:.3f} s | torch cpu ({torch.get_num_threads()} threads, {nsub} tensors): "
      f"{t2 - t1:.3f} s -> {(t2 - t1) * n / nsub:.1f} s per volume | max abs diff {err:.2e} (max entry {np.abs(b).max():.2e})")
