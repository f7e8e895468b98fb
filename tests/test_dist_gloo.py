"""world_size-2 checks of the multi-process host logic on CPU (gloo): rendezvous, unique-id
broadcast, max-over-ranks timing reduction, the row partition."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from flecsolve_b200 import dist as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(size))
    w = D.init(D.world_from_env())
    assert w.rank == rank and w.size == size and w.is_root == (rank == 0)
    payload = D.broadcast_bytes(w, bytes(range(128)) if w.is_root else None)
    t = D.max_over_ranks(w, 1.0 + rank)
    s = D.sum_over_ranks(w, 1.0 + rank)
    D.barrier(w)
    out.put((rank, payload == bytes(range(128)), t, s))
    D.finalize(w)


def test_two_rank_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res == [(0, True, 2.0, 3.0), (1, True, 2.0, 3.0)]


def test_equal_map_matches_oracle_partition():
    import oracle as O
    rp, col, val = O.stencil_csr(5, 7, 5)
    for colours in (1, 2, 3, 4, 6):
        assert np.array_equal(D.equal_map(35, colours), O.ParCSR(rp, col, val, colours=colours).partition())
    assert list(D.equal_map(10, 4)) == [0, 3, 6, 8, 10]


def test_single_rank_is_a_noop():
    w = D.init(D.world_from_env({}))
    assert w.size == 1 and D.max_over_ranks(w, 3.5) == 3.5 and D.broadcast_bytes(w, b"x") == b"x"
