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


class _TNet(nn.Module):
    """PointNet's input / feature transform regressor (k x k matrix, identity-initialised bias)."""

    def __init__(self, k):
        super().__init__()
        self.k = k
        self.conv1, self.conv2, self.conv3 = nn.Conv1d(k, 64, 1), nn.Conv1d(64, 128, 1), nn.Conv1d(128, 1024, 1)
        self.fc1, self.fc2, self.fc3 = nn.Linear(1024, 512), nn.Linear(512, 256), nn.Linear(256, k * k)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(1024)
        self.bn4, self.bn5 = nn.BatchNorm1d(512), nn.BatchNorm1d(256)

    def forward(self, x):
        B = x.shape[0]
        h = F.relu(self.bn1(self.conv1(x)))
        h = F.relu(self.bn2(self.conv2(h)))
        h = F.relu(self.bn3(self.conv3(h)))
        h = torch.max(h, 2)[0]
        h = F.relu(self.bn4(self.fc1(h)))
        h = F.relu(self.bn5(self.fc2(h)))
        h = self.fc3(h) + torch.eye(self.k, device=x.device).flatten()[None]
        return h.view(B, self.k, self.k)


class PointNetCls(nn.Module):
    """The standard PointNet classifier (Qi et al. 2017) with input and 64-d feature transforms -- the architecture of
    the reference's default victim (model/feature_models.py:71-98 + model/pointnet_utils.py:88-131), random-init,
    eval mode.  [B,3,K] -> logits [B,k].  Used as the victim of the attack-loop benchmark (the pretrained checkpoint
    is not available offline)."""

    def __init__(self, k=40, seed=0):
        super().__init__()
        torch.manual_seed(seed)
        self.stn, self.fstn = _TNet(3), _TNet(64)
        self.conv1, self.conv2, self.conv3 = nn.Conv1d(3, 64, 1), nn.Conv1d(64, 128, 1), nn.Conv1d(128, 1024, 1)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(1024)
        self.fc1, self.fc2, self.fc3 = nn.Linear(1024, 512), nn.Linear(512, 256), nn.Linear(256, k)
        self.bn4, self.bn5 = nn.BatchNorm1d(512), nn.BatchNorm1d(256)
        self.dropout = nn.Dropout(p=0.4)
        self.eval()

    def forward(self, x):
        trans = self.stn(x)
        h = torch.bmm(x.transpose(2, 1), trans).transpose(2, 1)
        h = F.relu(self.bn1(self.conv1(h)))
        tf = self.fstn(h)
        h = torch.bmm(h.transpose(2, 1), tf).transpose(2, 1)
        h = F.relu(self.bn2(self.conv2(h)))
        h = self.bn3(self.conv3(h))
        h = torch.max(h, 2)[0]
        h = F.relu(self.bn4(self.fc1(h)))
        h = F.relu(self.bn5(self.dropout(self.fc2(h))))
        return self.fc3(h)
