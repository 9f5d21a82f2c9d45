"""World-size-2 gloo test of the multi-GPU plumbing (frame sharding + record all-gather) on the CPU."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, q):
    sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
    from hn_b200 import parallel, runtime
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = parallel.shard_range(total, world, rank)
    per = (total + world - 1) // world
    # every frame's record is a function of its global index, so the gathered order can be checked
    idx = torch.arange(b, e, dtype=torch.float32)
    joints = idx[:, None, None] + torch.arange(63, dtype=torch.float32).reshape(1, 21, 3) / 100
    crops = torch.stack((idx, idx + 1, idx + 2, idx + 3), 1).to(torch.int64)
    has = (idx.to(torch.int64) % 2).to(torch.int32)
    rec = runtime.pack_records(joints, crops, has)
    allrec = parallel.gather_records(rec, per)
    if rank == 0:
        rows = []
        for r in range(world):
            rb, re_ = parallel.shard_range(total, world, r)
            rows.append(allrec[r * per: r * per + (re_ - rb)])
        j, c, h = runtime.unpack_records(torch.cat(rows))
        q.put((j[:, 0, 0].tolist(), c[:, 3].tolist(), h.tolist()))
    dist.destroy_process_group()


class _StubNet:
    """Stands in for HandNet on the CPU: a frame's record is a function of its content, so the test can tell which rank
    processed which frame.  Implements the two methods parallel.run_sharded needs."""

    def __init__(self, max_hands: int = 1):
        self.max_hands = max_hands

    def submit_records(self, images, depth, post):
        from hn_b200 import runtime
        idx = torch.stack([im.reshape(-1)[0] for im in images]) if images else torch.zeros(0)
        if self.max_hands > 1:     # one record row per (frame, hand) slot: slot h of frame i carries i + h / 10
            idx = (idx[:, None] + torch.arange(self.max_hands) / 10.0).reshape(-1)
            depth = depth.repeat_interleave(self.max_hands, dim=0)
            images = [None] * idx.numel()
        joints = idx[:, None, None] + torch.arange(63, dtype=torch.float32).reshape(1, 21, 3) / 100
        crops = torch.stack((idx, idx + 1, idx + 2, idx + 3), 1).to(torch.int64)
        has = (depth.reshape(len(images), -1)[:, 0] > 0.5).to(torch.int32)
        post(runtime.pack_records(joints, crops, has))
        return len(images)

    def result_records(self, ticket):
        return ticket


def _worker_run_sharded(rank, world, port, total, q):
    sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
    from hn_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the global batch: frame i is filled with the value i, its depth map with i % 2
    rgb = torch.arange(total, dtype=torch.float32).reshape(total, 1, 1, 1).expand(total, 3, 4, 4).contiguous()
    depth = (torch.arange(total) % 2).float().reshape(total, 1, 1, 1).expand(total, 1, 4, 4).contiguous()
    joints, crops, has = parallel.run_sharded(_StubNet(), rgb, depth)
    assert joints.shape == (total, 21, 3) and crops.shape == (total, 4) and has.shape == (total,)
    q.put((rank, joints[:, 0, 0].tolist(), crops[:, 3].tolist(), has.tolist()))
    dist.destroy_process_group()


def test_run_sharded_world2_gloo():
    """parallel.run_sharded (the product's shard -> run -> all-gather entry point) with an uneven split: every rank gets
    the whole batch's results in global frame order."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total, world, port = 7, 2, 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_run_sharded, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, j0, c3, h in got:
        assert j0 == [float(i) for i in range(total)]
        assert c3 == [i + 3 for i in range(total)]
        assert h == [bool(i % 2) for i in range(total)]


def _worker_run_sharded_hands(rank, world, port, total, hands, q):
    sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
    from hn_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rgb = torch.arange(total, dtype=torch.float32).reshape(total, 1, 1, 1).expand(total, 3, 4, 4).contiguous()
    depth = torch.ones(total, 1, 4, 4)
    joints, crops, has = parallel.run_sharded(_StubNet(max_hands=hands), rgb, depth)
    assert joints.shape == (total * hands, 21, 3) and has.shape == (total * hands,)
    q.put((rank, joints[:, 0, 0].tolist()))
    dist.destroy_process_group()


def test_run_sharded_world2_gloo_hand_slots():
    """HandNet(max_hands=H): H record rows per frame travel through the same all-gather; uneven split, padding rows of the
    short slice dropped, rows in (frame, hand) order on every rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total, world, hands, port = 5, 2, 3, 33000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_run_sharded_hands, args=(r, world, port, total, hands, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [round(i + h / 10.0, 4) for i in range(total) for h in range(hands)]
    for rank, j0 in got:
        assert [round(v, 4) for v in j0] == want


def test_run_sharded_single_process():
    sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
    from hn_b200 import parallel
    rgb = torch.arange(5, dtype=torch.float32).reshape(5, 1, 1, 1).expand(5, 3, 2, 2).contiguous()
    depth = torch.ones(5, 1, 2, 2)
    joints, crops, has = parallel.run_sharded(_StubNet(), rgb, depth)
    assert joints[:, 0, 0].tolist() == [0.0, 1.0, 2.0, 3.0, 4.0] and bool(has.all())


def test_shard_range_covers_everything():
    sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
    from hn_b200 import parallel
    for total in (1, 7, 8, 64, 255):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_record_allgather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total, world, port = 7, 2, 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    j0, c3, h = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert j0 == [float(i) for i in range(total)]
    assert c3 == [i + 3 for i in range(total)]
    assert h == [bool(i % 2) for i in range(total)]
