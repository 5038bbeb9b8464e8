"""Multi-GPU sharding logic (zshard.py): world_size-2 `gloo` runs on CPU for the host logic, and a 2-process run on
one GPU (gloo transport, so no kernel ever waits on another process) for the sharded ELBO / gradients."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _cpu_worker(rank, world, port, q):
    try:
        _init(rank, world, port)
        from draupnir_asr_b200.zshard import ShardContext
        torch.manual_seed(0)
        n, z = 7, 5                                                      # uneven leaf blocks: 4 + 3
        sh = ShardContext(rank, world, z, n)
        full = torch.randn(n, 3, dtype=torch.float64)
        w = torch.randn(n, 3, dtype=torch.float64)
        local = full[sh.leaf_lo:sh.leaf_hi].clone().requires_grad_(True)
        gathered = sh.gather_leaf_rows(local)
        assert torch.equal(gathered.detach(), full)
        # every rank's loss depends on the whole gathered tensor; the summed loss is sum_r (r+1) * <w, x^2>
        ((rank + 1) * (w * gathered ** 2).sum()).backward()
        want = sum(r + 1 for r in range(world)) * 2 * (w * full)[sh.leaf_lo:sh.leaf_hi]
        assert torch.allclose(local.grad, want, rtol=1e-14)
        # one all-reduce for gradients + loss
        p1 = torch.ones(4, dtype=torch.float64, requires_grad=True)
        p2 = torch.ones(2, 2, dtype=torch.float64, requires_grad=True)
        p1.grad = torch.full((4,), float(rank + 1), dtype=torch.float64)
        p2.grad = None if rank == 1 else torch.full((2, 2), 0.5, dtype=torch.float64)
        loss, failed = sh.reduce_step([p1, p2], torch.tensor(10.0 * (rank + 1), dtype=torch.float64),
                                      torch.tensor(rank, dtype=torch.int64))
        assert float(failed) == 1.0                                      # rank 1's failure is seen by both
        assert float(loss) == 30.0 and torch.equal(p1.grad, torch.full((4,), 3.0, dtype=torch.float64))
        assert torch.equal(p2.grad, torch.full((2, 2), 0.5, dtype=torch.float64))
        assert (sh.replicated_scale == 1.0) == (rank == 0)
        # z-slices of the conditional posterior (K6): uneven 3 + 2, gathered in order; rank 0's random numbers win
        zfull = torch.arange(z * 4, dtype=torch.float64).reshape(z, 4)
        assert torch.equal(sh.gather_z_rows(zfull[sh.z_lo:sh.z_hi].clone()), zfull)
        e = torch.full((3,), float(rank + 7), dtype=torch.float64)
        assert torch.equal(sh.broadcast_from_rank0(e), torch.full((3,), 7.0, dtype=torch.float64))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def _spawn(worker, world, *extra):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=worker, args=(r, world, port, q) + extra) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    return dict(res)


def test_split_even():
    from draupnir_asr_b200.zshard import split_even
    assert split_even(30, 2) == [(0, 15), (15, 30)]
    assert [hi - lo for lo, hi in split_even(30, 4)] == [8, 8, 7, 7]
    assert [hi - lo for lo, hi in split_even(30, 8)] == [4, 4, 4, 4, 4, 4, 3, 3]
    assert split_even(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]           # more ranks than items: empty tails


def test_gather_and_reduce_world2_gloo():
    res = _spawn(_cpu_worker, 2)
    assert res == {0: "ok", 1: "ok"}, res


# ------------------------------------------------------------------ GPU: sharded step == single-process step ----
def _gpu_worker(rank, world, port, q, guide):
    try:
        _init(rank, world, port)
        import __graft_entry__ as entry
        entry.build()
        from draupnir_asr_b200 import runtime, synthetic, zshard
        from oracle import draupnir_oracle as O
        n, L, z = 23, 31, 30
        problem = synthetic.make_problem(n, L, seed=2)
        init = synthetic.make_parameters(n, z_dim=z, seed=2)
        nets = O.Networks(z_dim=z, variational=(guide == "variational"), seed=2)
        state = {"decoder": nets.decoder.state_dict(), "embed": nets.embed.state_dict(), "h_0_MODEL": nets.h_0_MODEL}
        if guide == "variational":
            state.update(encoder=nets.encoder.state_dict(), embeddingencoder=nets.embeddingencoder.state_dict(),
                         h_0_GUIDE=nets.h_0_GUIDE)
        shard = zshard.ShardContext(rank, world, z, n)
        run = runtime.build_run(problem, guide, device="cuda:0", init=init, z_dim=z, shard=shard)
        runtime.load_networks(run, state)
        ref = O.OracleSVI(nets, init, guide, z_dim=z)
        eps = torch.randn(z, n, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
        import torch.distributions.normal as tn
        orig = tn._standard_normal
        tn._standard_normal = lambda shape, dtype, device: (eps.to(device=device, dtype=dtype).reshape(shape)
                                                            if torch.Size(shape).numel() == eps.numel()
                                                            else orig(shape, dtype=dtype, device=device))
        args = (problem.dataset_train, problem.patristic_matrix_train, problem.dataset_train_blosum, problem.blosum_weighted)
        for step in range(3):
            want = ref.step(*args, eps)
            got = run.step()
            assert abs(got - want) / abs(want) < 1e-9, (step, got, want)
        for (name, p_ref), (_, p) in zip(ref.nets.decoder.named_parameters(), run.Draupnir.decoder.named_parameters()):
            assert torch.allclose(p.detach().cpu(), p_ref.detach(), rtol=1e-6, atol=1e-9), name
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def _gpu_conditional_worker(rank, world, port, q):
    """z-sharded K6: each process eliminates the leaf block of its latent dimensions, all-gather of the means
    [z, ni] / of the draws; equal to the oracle's precision-partition arithmetic over all z."""
    try:
        _init(rank, world, port)
        import __graft_entry__ as entry
        entry.build()
        from draupnir_asr_b200 import runtime, synthetic, zshard
        from oracle import draupnir_oracle as O
        n, L, z = 45, 12, 30
        problem = synthetic.make_problem(n, L, seed=4)
        init = synthetic.make_parameters(n, z_dim=z, seed=4)
        shard = zshard.ShardContext(rank, world, z, n)
        run = runtime.build_run(problem, "delta_map", device="cuda:0", init=init, z_dim=z, shard=shard)
        calls = []
        from draupnir_asr_b200 import ops
        real = ops.ou_conditional
        ops.ou_conditional = lambda *a, **k: (calls.append(a[6].shape[0]), real(*a, **k))[1]
        est = {k: v.to("cuda:0") for k, v in init.items()}
        mean = run.Draupnir.conditional_samplingMAP(est, run.patristic_matrix_full)
        want = O.conditional_samplingMAP(init, problem.patristic_matrix_full, problem.internal_nodes)
        assert calls == [shard.z_hi - shard.z_lo], calls                  # only this rank's latent dimensions were eliminated
        assert torch.allclose(mean.cpu(), want, rtol=1e-8, atol=1e-10)
        eps = torch.randn(z, problem.n_internal, 2, dtype=torch.float64, generator=torch.Generator().manual_seed(9))
        draws = run.Draupnir.conditional_sampling_many(est, run.patristic_matrix_full, 2, eps=eps.to("cuda:0"))
        for s in range(2):
            w = O.conditional_sampling(init, problem.patristic_matrix_full, problem.internal_nodes, eps[:, :, s])
            assert torch.allclose(draws[s].cpu(), w, rtol=1e-7, atol=1e-8)
        torch.manual_seed(100 + rank)                                     # different generator states: rank 0's draw is used
        d2 = run.Draupnir.conditional_sampling_many(est, run.patristic_matrix_full, 1)
        d2 = d2.cpu()
        both = [torch.empty_like(d2) for _ in range(world)]
        dist.all_gather(both, d2)
        assert torch.equal(both[0], both[1])
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_conditional_two_processes_one_gpu(cuda):
    res = _spawn(_gpu_conditional_worker, 2)
    assert res == {0: "ok", 1: "ok"}, res


@pytest.mark.gpu
@pytest.mark.parametrize("guide", ["delta_map", "variational"])
def test_sharded_svi_equals_oracle_two_processes_one_gpu(cuda, guide):
    res = _spawn(_gpu_worker, 2, guide)
    assert res == {0: "ok", 1: "ok"}, res
