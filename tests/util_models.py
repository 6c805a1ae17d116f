"""Small victim networks for the attack-loop tests (any nn.Module mapping [B,3,K] -> logits works as a victim)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class TinyPointNet(nn.Module):
    """Shared-MLP + max-pool classifier, seeded init, eval mode (no BN/dropout state)."""

    def __init__(self, k=40, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.c1 = nn.Conv1d(3, 32, 1)
        self.c2 = nn.Conv1d(32, 64, 1)
        self.fc = nn.Linear(64, k)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
        self.eval()

    def forward(self, x):
        h = F.relu(self.c1(x))
        h = F.relu(self.c2(h))
        return self.fc(h.max(dim=2)[0])
