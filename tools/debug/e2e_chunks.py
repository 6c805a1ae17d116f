"""Config-5 shard through the host-buffer entry point: chunk size sweep (copies pipelined behind the kernels)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hit-adv_b200")]
import numpy as np, torch
from hitgeom.host import ChamferKnnHostStep

B, N = 1024, 16384
rng = np.random.default_rng(0)
ori = rng.standard_normal((B, N, 3)).astype(np.float32)
ori /= np.linalg.norm(ori, axis=-1).max(axis=1)[:, None, None]
adv = (ori + 0.01 * rng.standard_normal(ori.shape).astype(np.float32)).astype(np.float32)
ori_h, adv_h = torch.from_numpy(ori).pin_memory(), torch.from_numpy(adv).pin_memory()
grad_h = torch.empty_like(adv_h).pin_memory()
for chunk in (128, 256, 512):
    step = ChamferKnnHostStep(N, chunk_clouds=chunk)
    cl = np.empty(B, dtype=np.float32)
    for _ in range(2):
        step(adv_h.numpy(), ori_h.numpy(), grad_h.numpy())
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        loss, _ = step(adv_h.numpy(), ori_h.numpy(), grad_h.numpy())
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"chunk_clouds={chunk}: {np.median(ts):.2f} ms per step (wall), loss {loss:.6f}", flush=True)
    del step
