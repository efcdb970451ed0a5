"""world_size-2 gloo tests (CPU) of the pair-sharding plumbing used for N > 1 GPUs."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from roma_b200.sharding import shard_bounds


def test_shard_bounds():
    assert shard_bounds(64, 8) == [(8 * i, 8 * i + 8) for i in range(8)]
    assert shard_bounds(5, 2) == [(0, 3), (3, 5)]
    assert shard_bounds(1, 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]
    for n in range(0, 20):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


class _FakeModel:
    """Stands in for RegressionMatcher: deterministic per-pair function so that gathering can be checked."""
    h_resized = w_resized = 14
    upsample_res = (28, 28)
    upsample_preds = True
    symmetric = True

    def _get_device(self):
        return torch.device("cpu")

    def get_output_resolution(self):
        return self.upsample_res

    def match(self, a, b, im_A_high_res=None, im_B_high_res=None):
        n = a.shape[0]
        s = a.mean(dim=(1, 2, 3)) + 2 * b.mean(dim=(1, 2, 3)) + 3 * im_A_high_res.mean(dim=(1, 2, 3)) + 4 * im_B_high_res.mean(dim=(1, 2, 3))
        warp = s.view(n, 1, 1, 1).expand(n, 28, 56, 4).contiguous()
        cert = (s * 0.5).view(n, 1, 1).expand(n, 28, 56).contiguous()
        return warp, cert


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from roma_b200.sharding import match_sharded
    model = _FakeModel()
    g = torch.Generator().manual_seed(0)
    A, B = torch.randn(n_pairs, 3, 14, 14, generator=g), torch.randn(n_pairs, 3, 14, 14, generator=g)
    Ah, Bh = torch.randn(n_pairs, 3, 28, 28, generator=g), torch.randn(n_pairs, 3, 28, 28, generator=g)
    if rank == 0:
        res = match_sharded(model, A, B, Ah, Bh)
        ref_w, ref_c = model.match(A, B, Ah, Bh)
        q.put((torch.equal(res[0], ref_w), torch.equal(res[1], ref_c), tuple(res[0].shape)))
    else:
        assert match_sharded(model, None, None, n_pairs=n_pairs) is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [5, 1])
def test_match_sharded_gloo_world2(n_pairs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_pairs) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok_w, ok_c, shape = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok_w and ok_c and shape == (n_pairs, 28, 56, 4)
