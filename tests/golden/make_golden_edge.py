"""tests/golden/dgcnn_edge.npz: the UNMODIFIED reference `get_graph_feature` (model/dgcnn_cls.py:16-43) on this
container's CPU: edge features and the gradient autograd sends back to x for a seeded upstream gradient.
The function hard-codes `torch.device('cuda')` for its index base (dgcnn_cls.py:25); the harness hands the module
a `torch` proxy whose `device()` answers with the CPU device -- the reference file itself is untouched.
Build container only."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refload  # noqa: E402


class _TorchOnCpu:
    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*a, **k):
        return torch.device("cpu")


def main():
    ref = _refload.load()
    ref.dgcnn.torch = _TorchOnCpu()
    out = {}
    for tag, (B, C, N, k) in {"a": (2, 3, 128, 20), "b": (1, 64, 64, 20), "c": (1, 5, 100, 7)}.items():
        g = torch.Generator().manual_seed(70 + C)
        x = torch.randn(B, C, N, generator=g).requires_grad_()
        feat = ref.dgcnn.get_graph_feature(x, k=k)
        idx = ref.dgcnn.knn(x.detach(), k)
        w = torch.randn(feat.shape, generator=g)
        (feat * w).sum().backward()
        # (the upstream gradient w is not stored: the tests redraw it from the same generator, seed 70 + C, after x)
        out.update({f"{tag}_x": x.detach().numpy(), f"{tag}_idx": idx.numpy(), f"{tag}_feat": feat.detach().numpy(),
                    f"{tag}_grad": x.grad.numpy()})
        # explicit idx argument (AOF-style callers pass their own neighbour lists)
        feat2 = ref.dgcnn.get_graph_feature(x.detach(), k=k, idx=idx)
        assert torch.equal(feat2, feat.detach())
    np.savez_compressed(os.path.join(HERE, "dgcnn_edge.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("feat")})


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
