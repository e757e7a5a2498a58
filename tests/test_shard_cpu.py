"""Multi-process (gloo, world_size 2, CPU) checks of the data-parallel plumbing bench.py uses:
frame-pairs are sharded contiguously by rank with no data-path collective, the timing reduction is a
MAX over ranks, and only rank 0 reports."""
import os
import subprocess
import sys
import textwrap

import common  # noqa: F401

SCRIPT = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from d2t_b200 import parallel
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # 1. contiguous shard of the global pair list, both legs of a pair stay together
    lo, hi = parallel.shard_range(10, rank, world)
    owned = torch.zeros(10); owned[lo:hi] = 1
    dist.all_reduce(owned)
    assert torch.equal(owned, torch.ones(10)), owned          # every pair owned exactly once
    assert (hi - lo) in (5,)                                   # balanced
    # 2. shard-local roi image indices
    rois = torch.tensor([[6., 1, 2, 3, 4], [9., 1, 2, 3, 4]]) if rank == 1 else torch.tensor([[0., 1, 2, 3, 4]])
    loc = parallel.localize_rois(rois, lo)
    assert float(loc[:, 0].min()) >= 0 and float(loc[:, 0].max()) < hi - lo
    # 3. timing reduction: max over ranks, throughput = all units / max time
    t = parallel.max_over_ranks(torch.tensor([10.0 + rank, 20.0 - 5 * rank], dtype=torch.float64))
    assert t.tolist() == [11.0, 20.0], t
    v = parallel.throughput(units_per_rank=2, ms=float(t[0]), world=world)
    assert abs(v - 4 / 0.011) < 1e-6
    # 4. gather of per-shard results onto rank 0 in pair order (eval: results concatenated on the host)
    part = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1)
    full = parallel.gather_pairs(part, dst=0)
    if rank == 0:
        assert torch.equal(full.view(-1), torch.arange(10, dtype=torch.float32))
    else:
        assert full is None
    # 5. gradient all-reduce == gradient of the mean loss over the concatenated global batch
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    data = torch.randn(10, 8)
    target = torch.randn(10, 3)
    loss = torch.nn.functional.mse_loss(model(data[lo:hi]), target[lo:hi])
    loss.backward()
    nb = parallel.allreduce_gradients(model.parameters(), bucket_bytes=256)
    assert nb >= 2
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    ref.load_state_dict(model.state_dict())
    torch.nn.functional.mse_loss(ref(data), target).backward()
    for p, q in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-6)
    if rank == 0:
        print("OK", flush=True)
    dist.destroy_process_group()
""")


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "shard.py"
    script.write_text(SCRIPT % common.PKG)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK" in out.stdout
