"""Data-parallel sCM training step on two GPUs (skipped on a one-GPU box): per-stage NCCL all-reduce overlapped with the
backward + the replicated conditioning stage (training.GradientAllReduce) must leave, on EVERY rank, the mean of the two
ranks' gradients -- i.e. half the gradient a single process computes for the two samples as one batch."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from swift_b200 import synthetic as syn
    from swift_b200.training import GradientAllReduce
    from test_gpu_forward import build_net
    cfg = syn.SWIFT_SMALL
    net, _ = build_net(cfg)
    net = net.to(f"cuda:{rank}").train()
    lat, cond = syn.synthetic_fields(cfg, world, seed=5)
    x = torch.cat([lat, cond], 1)[rank:rank + 1].cuda().contiguous()
    t = torch.linspace(0.4, 1.3, world)[rank:rank + 1].cuda()
    aux = torch.full((1, 1), 0.6, device="cuda")
    cot = (torch.randn(world, cfg["out_channels"], *cfg["img_resolution"], generator=torch.Generator().manual_seed(2)) * 1e-5)[
        rank:rank + 1].cuda().contiguous()
    eng = net.model.train_engine()
    red = GradientAllReduce(net.model)
    eng.forward(x, None, t, aux)
    eng.backward(cot, on_stage=red.hook, cond_exchange=red.exchange_conditioning)
    red.finish()
    torch.cuda.synchronize()
    scales = {n: p for n, p in net.model.named_parameters() if n.endswith(".scale")}
    grads = {k: v.detach().cpu().clone() for k, v in eng.parameter_gradients(scales).items()}
    torch.save(grads, os.path.join(tmp, f"g{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_gradients_equal_half_the_two_sample_batch(tmp_path):
    import torch.multiprocessing as mp
    from swift_b200 import synthetic as syn
    from test_gpu_forward import build_net
    world, port = 2, 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0, g1 = (torch.load(os.path.join(tmp_path, f"g{r}.pt")) for r in range(world))
    cfg = syn.SWIFT_SMALL
    net, _ = build_net(cfg)
    net.train()
    lat, cond = syn.synthetic_fields(cfg, world, seed=5)
    x = torch.cat([lat, cond], 1).cuda().contiguous()
    t = torch.linspace(0.4, 1.3, world).cuda()
    aux = torch.full((world, 1), 0.6, device="cuda")
    cot = (torch.randn(world, cfg["out_channels"], *cfg["img_resolution"], generator=torch.Generator().manual_seed(2)) * 1e-5).cuda()
    eng = net.model.train_engine()
    eng.forward(x, None, t, aux)
    eng.backward(cot.contiguous())
    scales = {n: p for n, p in net.model.named_parameters() if n.endswith(".scale")}
    ref = {k: 0.5 * v.detach().cpu() for k, v in eng.parameter_gradients(scales).items()}
    for k in ref:
        assert torch.equal(g0[k], g1[k]), f"{k}: ranks disagree after the reduction"
        e = ((g0[k].double() - ref[k].double()).norm() / ref[k].double().norm().clamp_min(1e-30)).item()
        # same kernels, different summation split (per-rank then NCCL vs one batch): fp32 re-association only
        assert e < 2e-4, (k, e)
