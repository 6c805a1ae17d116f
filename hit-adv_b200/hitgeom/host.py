"""Host-buffer entry point: one CW-kNN distance step (ChamferkNNDist fwd+bwd, batch_avg=True) from host memory to host
memory through `hg_chamfer_knn_step_host_f32` -- chunks of clouds pipelined over two streams so the host<->device
copies hide behind the kernels (include/hitgeom.h).  No torch tensor crosses the boundary: arguments are numpy arrays
or CPU torch tensors (contiguous FP32; page-locked memory makes the copies asynchronous), the device buffers belong to
the session.

    step = ChamferKnnHostStep(points=16384, chunk_clouds=128)
    loss, cloud_loss = step(adv, ori, grad_out)          # adv, ori, grad_out: [B, points, 3] float32 host arrays
"""
import numpy as np
import torch

from ._lib import HitgeomError, check, lib

_METHODS = {"adv2ori": 0, "ori2adv": 1, "both": 2}


def _host_ptr(a, name, shape=None):
    if isinstance(a, torch.Tensor):
        if a.is_cuda or a.dtype != torch.float32 or not a.is_contiguous():
            raise HitgeomError(f"{name}: need a contiguous float32 CPU tensor")
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise HitgeomError(f"{name}: shape {tuple(a.shape)} != {tuple(shape)}")
        return a.data_ptr()
    if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
        raise HitgeomError(f"{name}: need a C-contiguous float32 numpy array or CPU tensor")
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise HitgeomError(f"{name}: shape {a.shape} != {tuple(shape)}")
    return a.ctypes.data


class ChamferKnnHostStep:
    """util/dist_utils.py:258-294 ChamferkNNDist(chamfer_method, knn_k, knn_alpha, chamfer_weight, knn_weight), forward
    and backward w.r.t. the adversarial clouds, on host buffers."""

    def __init__(self, points, chunk_clouds=128, chamfer_method="adv2ori", knn_k=5, knn_alpha=1.05, chamfer_weight=5.,
                 knn_weight=3.):
        if chamfer_method not in _METHODS:
            raise HitgeomError(f"chamfer_method must be one of {sorted(_METHODS)}")
        self.points, self.chunk = int(points), int(chunk_clouds)
        self.method, self.k, self.alpha = _METHODS[chamfer_method], int(knn_k), float(knn_alpha)
        self.w1, self.w2 = float(chamfer_weight), float(knn_weight)
        self._h = lib().hg_host_step_create(self.points, self.chunk, self.k)
        if not self._h:
            raise HitgeomError("hg_host_step_create failed: " + lib().hg_last_error().decode("utf-8", "replace"))

    def __call__(self, adv, ori, grad_out, weights=None, cloud_loss_out=None):
        B = adv.shape[0]
        shape = (B, self.points, 3)
        if cloud_loss_out is None:
            cloud_loss_out = np.empty(B, dtype=np.float32)
        loss = np.zeros(1, dtype=np.float32)
        check(lib().hg_chamfer_knn_step_host_f32(
            self._h, _host_ptr(adv, "adv", shape), _host_ptr(ori, "ori", shape), B, self.method, self.k, self.alpha,
            self.w1, self.w2, None if weights is None else _host_ptr(weights, "weights", (B,)), loss.ctypes.data,
            _host_ptr(cloud_loss_out, "cloud_loss_out", (B,)), _host_ptr(grad_out, "grad_out", shape)),
            "hg_chamfer_knn_step_host_f32")
        return float(loss[0]), cloud_loss_out

    def close(self):
        if getattr(self, "_h", None):
            lib().hg_host_step_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
