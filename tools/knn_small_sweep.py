"""Small-cloud self-kNN: whole call (seed pre-pass + main kernel) and main kernel alone, over the (QT, GP) variants,
with Z-order seeds (cold call) and temporal seeds (attack loop), against the streaming path it replaces
(development tool; the numbers quoted in DESIGN.md section 5.2 come from here).

    python tools/knn_small_sweep.py [--quick]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hit-adv_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from hitgeom import _lib  # noqa: E402
from hitgeom import functional as F  # noqa: E402
from util_inputs import clouds  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=10, before=None):
    for _ in range(3):
        if before:
            before()
        fn()
    ws, ks = [], []
    for _ in range(iters):
        if before:
            before()
        flush.zero_()
        torch.cuda.synchronize()
        _lib.prof_enable(True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        kms, n = _lib.prof_read("knn")
        _lib.prof_enable(False)
        ws.append(a.elapsed_time(b))
        ks.append(kms / max(n, 1))
    return sorted(ws)[iters // 2], sorted(ks)[iters // 2]


quick = "--quick" in sys.argv
shapes = [(388, 1024, 6, "gauss"), (388, 1024, 6, "surface"), (32, 1024, 20, "gauss"), (388, 1024, 20, "gauss"),
          (64, 2000, 6, "gauss"), (1024, 512, 6, "gauss")]
if quick:
    shapes = shapes[:1]
for (B, N, k1, kind) in shapes:
    x = torch.from_numpy(clouds(B, N, 1234, kind)).cuda()
    pairs = float(B) * N * N
    F.tune_knn_small(-1)
    F.force_knn_shape(0, 0)
    w, k = timed(lambda: F.knn_self(x, k1))
    ref = F.knn_self(x, k1)
    print(f"B={B} N={N} k1={k1} {kind}: streaming path     whole {w * 1e3:7.1f} us  kernel {k * 1e3:7.1f} us  "
          f"{pairs / k * 1e3:.3e} pair-evals/s", flush=True)
    F.tune_knn_small(0)
    qts = (1, 2, 4) if k1 <= 6 else (1, 2) if k1 <= 20 else (1,)
    for qt in qts:
        for gp in (1, 2, 4):
            F.force_knn_shape(qt, gp)
            try:
                w, k = timed(lambda: F.knn_self(x, k1))
                out = F.knn_self(x, k1)
                same = torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
                line = (f"   small QT={qt} GP={gp}: cold whole {w * 1e3:7.1f} us  kernel {k * 1e3:7.1f} us  "
                        f"{pairs / k * 1e3:.3e}/s {'ok' if same else 'MISMATCH'} |")
                state = torch.empty((B, N, k1), dtype=torch.int32, device="cuda")
                for tag, step in (("same", 0.0), ("1e-3", 1e-3), ("5e-3", 5e-3)):
                    xm = x + step * torch.randn_like(x)

                    def refresh():  # the state holds x's neighbours before every timed call on the moved cloud
                        F.knn_self(x, k1, state=state, state_valid=False)

                    w, k = timed(lambda: F.knn_self(xm, k1, state=state, state_valid=True), iters=6, before=refresh)
                    line += f" temporal({tag}) whole {w * 1e3:6.1f} kernel {k * 1e3:6.1f} |"
                print(line, flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"   small QT={qt} GP={gp}: {e}", flush=True)
    F.force_knn_shape(0, 0)
