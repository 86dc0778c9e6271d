"""CPU test of the N>1 host logic: world_size-2 gloo process group, frame sharding without any data-path collective,
whole-job throughput = frames of all ranks / max over ranks of the time."""
import os
import socket
import sys

import pytest

from conftest import ROOT, load_binding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch.distributed as dist

    pkg = load_binding()
    from elas_b200 import sharding

    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = sharding.Group(dist)
    first, last = sharding.frame_range(rank, world, 4)
    # every rank generates ITS frames only (host-side generator; no GPU involved)
    digest = []
    for f in range(first, last):
        L, R = pkg.binding.synth_pair(f, 64, 32)
        digest.append(int(L.astype(np.int64).sum()))
    ms = 100.0 * (rank + 1)  # rank 1 is the slow one
    value = g.throughput(last - first, ms)
    g.barrier()
    q.put((rank, first, last, digest, value, g.max(ms), g.sum(last - first)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_aggregation():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, l0, d0, v0, m0, s0), (r1, f1, l1, d1, v1, m1, s1) = res
    assert (f0, l0, f1, l1) == (0, 4, 4, 8)  # contiguous, disjoint
    assert d0 != d1  # different frames on different ranks
    assert m0 == m1 == 200.0 and s0 == s1 == 8.0
    assert v0 == v1 == pytest.approx(8 / 0.2)


def test_split_frames_covers_everything():
    pkg = load_binding()
    from elas_b200 import sharding

    for n, w in ((1024, 8), (10, 3), (2, 4), (0, 2)):
        parts = sharding.split_frames(n, w)
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.frame_range(2, 2, 4)
