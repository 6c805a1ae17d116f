"""CW-kNN attack loops -- the callers of the distance kernels, behind the reference's
own attacker interface:

    CWKNN (model, adv_func, dist_func, clip_func, attack_lr=1e-3, num_iter=2500).attack(data, target)   CW/kNN.py:14-151
    CWUKNN(model, adv_func, dist_func, clip_func, attack_lr=1e-3, num_iter=2500, pre_head=None)         CW/UKNN.py:14-159
        data [B,K,3] (or [B,K,6] with normals), target [B] -> (np.ndarray [B,K,3] float32, success count)

Same algorithm and the same random draw (`torch.randn((B,3,K))` from the CPU generator, kNN.py:61-62), the same
optimiser (Adam on the channel-first cloud), `dist_loss = dist_func(adv^T, ori^T).mean() * K`, clip after every step.
What changes:
  * `dist_func` is hitgeom's ChamferkNNDist / ChamferDist / ... (one fused pass per term);
  * the original cloud is transposed ONCE, not on every iteration (kNN.py:104-106 re-materialises it 2500 times);
  * nothing inside the loop reads a value back: the per-iteration `(pred == target).sum().item()` (kNN.py:90) and
    the timers' implicit syncs are gone -- the success count is read at the `num_iter // 5` progress marks only if
    `verbose`, and once at the end; `torch.cuda.empty_cache()` every 100 iterations (kNN.py:131) is dropped;
  * so the iteration is sync-free and, with `graph=True`, everything after the first iterations is replayed as ONE
    CUDA graph launch per step (victim forward/backward, loss kernels, Adam, clip).
"""
import numpy as np
import torch
import torch.optim as optim


class CWKNN:
    untargeted = False

    def __init__(self, model, adv_func, dist_func, clip_func, attack_lr=1e-3, num_iter=2500, verbose=False,
                 graph=False, capturable_adam=None):
        self.model = model.cuda()
        self.model.eval()
        self.adv_func = adv_func
        self.dist_func = dist_func
        self.clip_func = clip_func
        self.attack_lr = attack_lr
        self.num_iter = num_iter
        self.verbose = verbose
        self.graph = graph
        # Adam with its step counter on the device (needed for capture; bias corrections then round in FP32 on the
        # device instead of in Python doubles -- a 1e-7 relative difference in the step size)
        self.capturable_adam = graph if capturable_adam is None else capturable_adam
        self.loop_ms = 0.0  # device time of the iteration loop of the last attack() (for the benchmark)
        self.replay_ms, self.replays = 0.0, 0  # of which: the graph-replayed iterations

    # -- pieces ------------------------------------------------------------------------------------------------
    def _logits(self, adv_data):
        logits = self.model(adv_data)
        return logits[0] if isinstance(logits, tuple) else logits  # PointNet returns (logits, trans, trans_feat)

    def _clip(self, adv_data, ori_data, normal):
        return self.clip_func(adv_data, ori_data)

    def _success(self, pred, target):
        return (pred != target) if self.untargeted else (pred == target)

    def _step(self, adv_data, ori_data, ori_t, normal, target, opt, stats):
        K = adv_data.shape[2]
        logits = self._logits(adv_data)
        adv_loss = self.adv_func(logits, target).mean()
        dist_loss = self.dist_func(adv_data.transpose(1, 2).contiguous(), ori_t).mean() * K
        loss = adv_loss + dist_loss
        opt.zero_grad(set_to_none=False)
        loss.backward()
        opt.step()
        if self.clip_func is not None:
            with torch.no_grad():
                adv_data.copy_(self._clip(adv_data.detach(), ori_data, normal))
        with torch.no_grad():  # device-side progress record (no sync): [success count, adv_loss, dist_loss]
            stats[0] = self._success(torch.argmax(logits, dim=1), target).sum()
            stats[1] = adv_loss
            stats[2] = dist_loss

    # -- the attack (CW/kNN.py:40-151) -------------------------------------------------------------------------------
    def attack(self, data, target):
        B, K = data.shape[:2]
        data = data.float().cuda().detach().transpose(1, 2).contiguous()
        ori_data = data.clone().detach()
        if ori_data.shape[1] == 3:
            normal = None
        else:
            normal = ori_data[:, 3:, :].contiguous()
            ori_data = ori_data[:, :3, :].contiguous()
        target = target.long().cuda().detach()
        ori_t = ori_data.transpose(1, 2).contiguous()  # loop-invariant

        adv_data = ori_data.clone().detach() + torch.randn((B, 3, K)).cuda() * 1e-7
        adv_data.requires_grad_()
        # the kNN term sees a cloud that moves by a learning-rate step per iteration: seed its thresholds from the
        # previous iteration's neighbours (same results, no spatial pre-pass; hitgeom.dist_utils.KNNDist.temporal_seeds)
        if hasattr(self.dist_func, "temporal_seeds"):
            self.dist_func.temporal_seeds(True)
        opt = optim.Adam([adv_data], lr=self.attack_lr, weight_decay=0., capturable=self.capturable_adam)
        stats = torch.zeros(3, device=adv_data.device)
        marks = max(self.num_iter // 5, 1)

        def report(it):
            s = stats.tolist()  # the only read-back inside the loop, and only when asked for
            print('Iteration {}/{}, success {}/{}\nadv_loss: {:.4f}, dist_loss: {:.4f}'.format(
                it, self.num_iter, int(s[0]), B, s[1], s[2]))

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eg = torch.cuda.Event(enable_timing=True)
        args = (adv_data, ori_data, ori_t, normal, target, opt, stats)
        graph, warm = None, min(3, self.num_iter)
        e0.record()
        for iteration in range(self.num_iter):
            if self.graph and iteration == warm:
                # warm-up iterations ran eagerly (allocator pools, Adam state); capture one iteration and replay it
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(graph, stream=side):
                        self._step(*args)
                torch.cuda.current_stream().wait_stream(side)
                eg.record()
            if graph is not None:  # (capture records the iteration without running it)
                graph.replay()
            else:
                self._step(*args)
            if self.verbose and iteration % marks == 0:
                report(iteration)
        e1.record()

        with torch.no_grad():
            pred = torch.argmax(self._logits(adv_data), dim=-1)
            success_num = int(self._success(pred, target).sum().item())
        self.loop_ms = e0.elapsed_time(e1)
        self.replays = max(self.num_iter - warm, 0) if graph is not None else 0
        self.replay_ms = eg.elapsed_time(e1) if graph is not None else 0.0
        if self.verbose:
            print('Successfully attack {}/{}'.format(success_num, B))
        return adv_data.detach().transpose(1, 2).contiguous().cpu().numpy(), success_num


class CWUKNN(CWKNN):
    """Untargeted variant (CW/UKNN.py): success = prediction differs from the label, optional `pre_head` module in
    front of the victim (UKNN.py:82-85), and the clip function receives the normals (UKNN.py:121-122)."""
    untargeted = True

    def __init__(self, model, adv_func, dist_func, clip_func, attack_lr=1e-3, num_iter=2500, pre_head=None,
                 verbose=False, graph=False, capturable_adam=None):
        super().__init__(model, adv_func, dist_func, clip_func, attack_lr, num_iter, verbose, graph, capturable_adam)
        self.pre_head = pre_head

    def _logits(self, adv_data):
        logits = self.model(self.pre_head(adv_data)) if self.pre_head is not None else self.model(adv_data)
        return logits[0] if isinstance(logits, tuple) else logits

    def _clip(self, adv_data, ori_data, normal):
        return self.clip_func(adv_data, ori_data, normal)
