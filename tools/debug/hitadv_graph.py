import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import bench
from hitgeom.hit_adv import HiT_ADV, UntargetedLogitsAdvLoss
from util_models import PointNetCls, TinyPointNet
data, target = bench.hitadv_inputs(16, 1024, 1)
for mk in (TinyPointNet, PointNetCls):
    model = mk(40, seed=0).cuda()
    for hp in (dict(bench.HITADV_HP), dict(bench.HITADV_HP, cd_weight=0), dict(bench.HITADV_HP, cd_weight=0, hide_weight=0), dict(bench.HITADV_HP, cd_weight=0, hide_weight=0, ker_weight=0)):
        try:
            att = HiT_ADV(model, UntargetedLogitsAdvLoss(kappa=30.0), clip_func=None, binary_step=1, num_iter=8, graph=True, **hp)
            torch.manual_seed(0)
            att.attack_device(data, target)
            torch.cuda.synchronize()
            print(mk.__name__, {k: hp[k] for k in ("cd_weight", "hide_weight", "ker_weight")}, "OK", att.replay_ms, flush=True)
        except Exception as e:
            print(mk.__name__, {k: hp[k] for k in ("cd_weight", "hide_weight", "ker_weight")}, "FAILED", flush=True)
            traceback.print_exc(limit=6)
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
