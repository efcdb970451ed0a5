"""NCCL world-size-2 test of the multi-GPU path: `sharding.match_sharded` (rank 0 holds the batch, inputs scattered and results
gathered over NCCL) must equal a single-GPU `match()` of the same pairs.  Needs two GPUs (skipped otherwise; run with
`gpurun --gpus 2 -- python -m pytest tests/test_sharding_gpu.py -m gpu`)."""
import os

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n_pairs, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from roma_b200 import roma_outdoor, sharding, synthetic
    mw, dw = synthetic.make_weights(0)
    model = roma_outdoor(dev, weights=mw, dinov2_weights=dw, coarse_res=112, upsample_res=168, amp_dtype=torch.float32)
    if rank == 0:
        A, B, Ah, Bh = (t.to(dev) for t in synthetic.make_pair(n_pairs, 112, 168, seed=3))
        seen = []
        res = sharding.match_sharded(model, A, B, Ah, Bh, max_batch=2, on_batch=lambda w, c: seen.append(w.shape[0]))
        ref_w, ref_c = model.match(A, B, im_A_high_res=Ah, im_B_high_res=Bh)
        dw_, dc_ = (res[0] - ref_w).abs().max().item(), (res[1] - ref_c).abs().max().item()
        q.put((dw_, dc_, tuple(res[0].shape), seen))
    else:
        assert sharding.match_sharded(model, None, None, n_pairs=n_pairs, max_batch=2) is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [5, 2])
def test_match_sharded_nccl_world2(n_pairs):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + n_pairs) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    dw, dc, shape, seen = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    assert shape == (n_pairs, 168, 336, 4)
    assert dw <= 1e-6 and dc <= 1e-6, (dw, dc)          # per-pair arithmetic does not depend on the batch a pair travels in
    assert sum(seen) == (n_pairs + 1) // 2 and max(seen) <= 2
