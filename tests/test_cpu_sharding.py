"""Multi-process test of the instance-sharding driver on CPU: world_size 2, gloo backend."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "hit-adv_b200"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from hitgeom import sharding

    sharding.init(backend="gloo")
    data = torch.arange(n * 4 * 3, dtype=torch.float32).view(n, 4, 3)
    labels = torch.arange(n)

    def attack(local, lab):  # stands in for an attack on the rank's block of instances
        return local * 2 + lab[:, None, None].float(), {"at_num": float(local.shape[0]), "batches": 1.0}

    out, counters = sharding.run_sharded(attack, data, labels)
    expect = data * 2 + labels[:, None, None].float()
    ok = torch.equal(out, expect) and counters == {"at_num": float(n), "batches": float(world)}
    t = sharding.max_over_ranks(10.0 + rank)
    lo, hi = sharding.shard_range(n)
    q.put((rank, ok, t, lo, hi))
    sharding.barrier()
    torch.distributed.destroy_process_group()


def test_run_sharded_world2_gloo():
    n, world = 7, 2  # ragged: 4 + 3
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [11.0, 11.0]
    assert [(r[3], r[4]) for r in res] == [(0, 4), (4, 7)]


def test_shard_range_covers_everything():
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "hit-adv_b200"))
    from hitgeom import sharding

    for n in (0, 1, 7, 388, 8192):
        for w in (1, 2, 4, 8):
            r = [sharding.shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
