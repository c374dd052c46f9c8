"""CPU, world_size 2 and 4 over gloo: the host side of the gallery-sharded search -- row
partition, global ids, the single all-gather payload, the merge contract.  The two device
steps are injected (the oracle stands in for them here; on the GPU they are ofx_topk_search /
ofx_topk_merge, covered by tests/test_gpu_search.py::test_sharded_search_equals_single)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restatement as R
from outfitx_b200 import synth
from outfitx_b200.search import ShardedSearch, shard_rows

N, NQ, K = 5003, 24, 10


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _CpuShard:
    def __init__(self, rows, lo):
        self.rows, self.id_offset = rows, lo


def _oracle_local(queries, gallery, k, metric, exact):
    i, s = R.search(queries.numpy(), gallery.rows, k=k, metric=metric, id_offset=gallery.id_offset)
    return torch.from_numpy(i), torch.from_numpy(s)


def _oracle_merge(idx, score, k):
    w, nq, kk = idx.shape
    i = idx.permute(1, 0, 2).reshape(nq, w * kk).numpy()
    s = score.permute(1, 0, 2).reshape(nq, w * kk).numpy()
    s = np.where(i < 0, -np.inf, s)
    i = np.where(i < 0, np.iinfo(np.int64).max, i)      # padding ranks last
    mi, ms = R.merge_topk(i, s, k)
    mi = np.where(np.isneginf(ms), -1, mi)
    return torch.from_numpy(mi), torch.from_numpy(ms)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gal = synth.make_items(N, 512, seed=5, dup=200)
        q = torch.from_numpy(synth.make_queries(NQ, 1024, seed=6))
        lo, hi = shard_rows(N, rank, world)
        shard = _CpuShard(gal[lo:hi], lo)
        idx, score = ShardedSearch(local=_oracle_local, merge=_oracle_merge).search(q, shard, K, "l2", True)
        if rank == 0:
            np.savez(out, idx=idx.numpy(), score=score.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_search_host_logic(tmp_path, world):
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    gal = synth.make_items(N, 512, seed=5, dup=200)
    want_i, want_s = R.search(synth.make_queries(NQ, 1024, seed=6), gal, k=K)
    assert np.array_equal(got["idx"], want_i)
    assert np.array_equal(got["score"], want_s)


def test_tiny_shards_pad_with_minus_one():
    """More ranks than rows: empty shards contribute only padding (-1 / -inf)."""
    gal = synth.make_items(3, 512, seed=1)
    q = synth.make_queries(2, 1024, seed=2)
    parts = []
    for r in range(4):
        lo, hi = shard_rows(3, r, 4)
        parts.append(R.search(q, gal[lo:hi], k=5, id_offset=lo))
    idx = torch.from_numpy(np.stack([p[0] for p in parts]))
    score = torch.from_numpy(np.stack([p[1] for p in parts]))
    mi, ms = _oracle_merge(idx, score, 5)
    want_i, want_s = R.search(q, gal, k=5)
    assert np.array_equal(mi.numpy(), want_i) and np.array_equal(ms.numpy(), want_s)
