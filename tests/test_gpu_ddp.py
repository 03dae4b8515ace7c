"""Data-parallel sCM training step on two GPUs (skipped on a one-GPU box): per-stage NCCL all-reduce overlapped with the
backward + the replicated conditioning stage (training.GradientAllReduce) must leave, on EVERY rank, the mean of the two
ranks' gradients -- i.e. half the gradient a single process computes for the two samples as one batch."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(cfg, logvar):
    from swift_b200 import synthetic as syn
    from swift_b200.precond import PassPrecond
    from test_gpu_forward import build_net
    if not logvar:
        return build_net(cfg)[0]
    model_cfg = dict(_target_="swift_b200.swinv2.SwinV2", window_size=cfg["window_size"], shift_size=cfg["shift_size"],
                     patch_size=cfg["patch_size"], depth=cfg["depth"], dim=cfg["dim"], heads=cfg["heads"], logvar=True,
                     timestep_weight=1.0)
    n_img = cfg["out_channels"]
    net = PassPrecond(model_cfg, img_resolution=cfg["img_resolution"], img_channels=n_img,
                      condition_channels=cfg["in_channels"] - n_img, auxiliary_dim=cfg["auxiliary_dim"], sigma_data=1.0)
    net.load_state_dict(syn.random_state_dict(cfg, seed=1, prefix="model.", logvar=True), strict=True)
    return net.cuda().eval()


def _worker(rank, world, port, tmp, logvar):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from swift_b200 import synthetic as syn
    from swift_b200.training import GradientAllReduce
    cfg = syn.SWIFT_SMALL
    net = _build(cfg, logvar).to(f"cuda:{rank}").train()
    lat, cond = syn.synthetic_fields(cfg, world, seed=5)
    x = torch.cat([lat, cond], 1)[rank:rank + 1].cuda().contiguous()
    t = torch.linspace(0.4, 1.3, world)[rank:rank + 1].cuda()
    aux = torch.full((1, 1), 0.6, device="cuda")
    cot = (torch.randn(world, cfg["out_channels"], *cfg["img_resolution"], generator=torch.Generator().manual_seed(2)) * 1e-5)[
        rank:rank + 1].cuda().contiguous()
    eng = net.model.train_engine()
    red = GradientAllReduce(net.model)
    eng.forward(x, None, t, aux)
    dlv = torch.tensor([0.7, -1.3])[rank:rank + 1].cuda() if logvar else None
    eng.backward(cot, on_stage=red.hook, cond_exchange=red.exchange_conditioning, dlogvar=dlv)
    red.finish()
    torch.cuda.synchronize()
    scales = {n: p for n, p in net.model.named_parameters() if n.endswith(".scale")}
    grads = {k: v.detach().cpu().clone() for k, v in eng.parameter_gradients(scales).items()}
    torch.save(grads, os.path.join(tmp, f"g{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("logvar", [False, True], ids=["plain", "logvar_head"])
def test_two_rank_gradients_equal_half_the_two_sample_batch(tmp_path, logvar):
    """``logvar_head``: the model carries logvar_embed and dL/dlogvar (one value per sample) is gathered with the other inputs of
    the replicated conditioning stage."""
    import torch.multiprocessing as mp
    from swift_b200 import synthetic as syn
    world, port = 2, 29600 + os.getpid() % 300 + (17 if logvar else 0)
    mp.spawn(_worker, args=(world, port, str(tmp_path), logvar), nprocs=world, join=True)
    g0, g1 = (torch.load(os.path.join(tmp_path, f"g{r}.pt")) for r in range(world))
    cfg = syn.SWIFT_SMALL
    net = _build(cfg, logvar)
    net.train()
    lat, cond = syn.synthetic_fields(cfg, world, seed=5)
    x = torch.cat([lat, cond], 1).cuda().contiguous()
    t = torch.linspace(0.4, 1.3, world).cuda()
    aux = torch.full((world, 1), 0.6, device="cuda")
    cot = (torch.randn(world, cfg["out_channels"], *cfg["img_resolution"], generator=torch.Generator().manual_seed(2)) * 1e-5).cuda()
    eng = net.model.train_engine()
    eng.forward(x, None, t, aux)
    eng.backward(cot.contiguous(), dlogvar=torch.tensor([0.7, -1.3]).cuda() if logvar else None)
    scales = {n: p for n, p in net.model.named_parameters() if n.endswith(".scale")}
    ref = {k: 0.5 * v.detach().cpu() for k, v in eng.parameter_gradients(scales).items()}
    for k in ref:
        assert torch.equal(g0[k], g1[k]), f"{k}: ranks disagree after the reduction"
        e = ((g0[k].double() - ref[k].double()).norm() / ref[k].double().norm().clamp_min(1e-30)).item()
        # same kernels, different summation split (per-rank then NCCL vs one batch): fp32 re-association only
        assert e < 2e-4, (k, e)
