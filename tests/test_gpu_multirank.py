"""Data-parallel parity on real GPUs (needs >= 2 visible B200s; skipped otherwise -- run it with `gpurun --gpus 2`):

  * N ranks on shards of a global batch == one rank on the whole batch.  With
    `--norm_G spectralspadeinstance3x3` every normalisation is per sample, so the averaged gradient of the shards IS the
    gradient of the full batch; checked on the all-reduced G+E gradients of one generator step and on the losses.
  * replicas start identical without any help from the caller (the trainer broadcasts rank 0's state), checkpoints are
    written by rank 0 only;
  * the multi-rank CUDA-graph path ([fwd+bwd] graph -> eager NCCL all-reduce -> [Adam] graph) reproduces the eager
    multi-rank path."""
import os
import socket
from types import SimpleNamespace

import pytest
import torch
import torch.multiprocessing as mp

from oracle import seg2eye_oracle as O

pytestmark = pytest.mark.gpu
WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _opts(tmp):
    o = O.make_opt(ngf=16, ndf=16, lambda_l1=10.0, norm_G="spectralspadeinstance3x3")
    d = vars(o).copy()
    d.update(gpu_ids=[0], init_type="xavier", init_variance=0.02, netD_subarch="n_layer", continue_train=False,
             which_epoch="latest", checkpoints_dir=tmp, name="mr", no_vgg_loss=True, lambda_openeds=0.0,
             lambda_style_w=0.0, lambda_style_feat=0.0, lambda_gram=0.0, netG="spadestyle", netD="multiscale")
    return o, SimpleNamespace(**d)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _worker(rank, port, tmp, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(WORLD), RANK=str(rank), LOCAL_RANK=str(rank))
        torch.cuda.set_device(rank)
        import datetime
        import torch.distributed as dist
        # a short collective timeout: a desynchronised rank must fail the test, not hang the box
        dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=90), device_id=torch.device("cuda", rank))
        from seg2eye_b200 import parallel
        from seg2eye_b200.trainers.pix2pix_trainer import Pix2PixTrainer
        oopt, opt = _opts(tmp)
        opt.gpu_ids = [rank]
        torch.manual_seed(100 + rank)                      # different random init per rank on purpose
        tr = Pix2PixTrainer(opt)                           # joins the process group and broadcasts rank 0's state
        assert tr.world == WORLD and dist.is_initialized()
        m = tr.pix2pix_model
        res = {}
        # (1) replicas identical after construction
        probe = torch.cat([p.detach().reshape(-1)[:64] for p in list(m.netG.parameters())[:8]] +
                          [b.detach().float().reshape(-1)[:64] for b in list(m.netG.buffers())[:8]])
        both = [torch.empty_like(probe) for _ in range(WORLD)]
        dist.all_gather(both, probe)
        res["replicas_identical"] = bool(torch.equal(both[0], both[1]))
        # deterministic weights for the gradient comparison
        sds = dict(G=O.synth_state(O.generator_shapes(oopt), 71), D=O.synth_state(O.discriminator_shapes(oopt), 72),
                   E=O.synth_state(O.encoder_shapes(oopt), 73))
        for net, k in ((m.netG, "G"), (m.netD, "D"), (m.netE, "E")):
            net.load_state_dict({a: b.clone() for a, b in sds[k].items()})
        full = O.synth_batch(oopt, 2 * WORLD, 74)
        shard = {k: v[2 * rank:2 * rank + 2].clone() for k, v in full.items()}
        names = ["fc.weight", "head_0.conv_0.weight_orig", "up_1.norm_0.spade.mlp_gamma.weight", "up_3.conv_1.weight_orig",
                 "up_3.norm_1.adain.linear.weight", "conv_img.weight"]
        pG = dict(m.netG.named_parameters())
        m.train()

        def reload():
            for net, k in ((m.netG, "G"), (m.netD, "D"), (m.netE, "E")):
                net.load_state_dict({a: b.clone() for a, b in sds[k].items()})
            for o_ in (tr.optimizer_G, tr.optimizer_D):       # fresh Adam state
                o_.state.clear()
                for g in o_.param_groups:
                    g.pop('_s2e_state', None)

        def two_iterations():
            out = []
            for _ in range(2):
                d = {k: v.clone() for k, v in shard.items()}
                tr.run_generator_one_step(d)
                tr.run_discriminator_one_step(d)
                out.append(torch.stack([v.reshape(-1)[0].detach().float() for v in tr.get_latest_losses().values()]).cpu())
            return out

        # (2) multi-rank CUDA graphs ([fwd+bwd] graph -> NCCL all-reduce -> [Adam] graph) == multi-rank eager
        tr.enable_cuda_graphs({k: v.cuda() for k, v in shard.items()}, warmup=1)
        graph = two_iterations()
        tr.disable_cuda_graphs()
        reload()
        eager = two_iterations()
        res["graph_vs_eager"] = [float((a - b).abs().max()) for a, b in zip(eager, graph)]
        res["eager_losses"] = eager[1]
        tr.g_losses = tr.d_losses = None
        tr.generated = None
        reload()
        # (3a) rank 0 alone: one generator step on the WHOLE batch with the collective switched off
        if rank == 0:
            tr.reducer_G.remove_hooks()
            tr.optimizer_G.zero_grad()
            losses_f, _ = tr._fb('G', {k: v.clone() for k, v in full.items()})
            g_full = {n: pG[n].grad.detach().clone() for n in names}
            res["losses_full"] = torch.stack([v.reshape(-1)[0].detach() for v in losses_f.values()]).cpu()
            del losses_f
            for net, k in ((m.netG, "G"), (m.netD, "D"), (m.netE, "E")):      # u / v advanced: start again from the same state
                net.load_state_dict({a: b.clone() for a, b in sds[k].items()})
        else:
            tr.reducer_G.remove_hooks()
        dist.barrier()
        # (3b) both ranks: sharded gradients, averaged across ranks
        tr.optimizer_G.zero_grad()
        tr.reducer_G.armed, tr.reducer_D.armed = True, False
        losses, _ = tr._fb('G', dict(shard))
        tr.reducer_G.allreduce()
        lsum = torch.stack([v.reshape(-1)[0].detach() for v in losses.values()])
        dist.all_reduce(lsum)
        res["losses_sharded_mean"] = (lsum / WORLD).cpu()
        del losses
        if rank == 0:
            res["grad_err"] = {n: rel(pG[n].grad, g_full[n]) for n in names}
        dist.barrier()
        # (4) rank-0-only checkpoint
        tr.save("latest")
        res["ckpt_exists"] = os.path.exists(os.path.join(tmp, "mr", "latest_net_G.pth"))
        q.put((rank, res, None))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:      # noqa: BLE001 -- report to the parent instead of hanging it
        import traceback
        q.put((rank, None, traceback.format_exc()))


@pytest.mark.skipif(torch.cuda.device_count() < WORLD, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_two_ranks_match_one_rank_and_graph_path(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, str(tmp_path), q), daemon=True) for r in range(WORLD)]
    for p in procs:
        p.start()
    got = []
    try:
        for _ in procs:
            got.append(q.get(timeout=300))
            if got[-1][2] is not None:      # a rank failed: its peers are stuck in a collective, do not wait for them
                break
    finally:
        for p in procs:
            p.join(5 if (got and got[-1][2] is not None) else 60)
            if p.is_alive():
                p.kill()
    for rank, res, err in got:
        assert err is None, "rank %d: %s" % (rank, err)
    assert len(got) == WORLD
    res = {rank: r for rank, r, _ in got}
    for r in res.values():
        assert r["replicas_identical"] and r["ckpt_exists"]
    r0 = res[0]
    # sharded == full batch: gradients within the bf16 noise of two different batch compositions (the 10x8 / 20x16 maps are
    # tiled across samples, so a batch of 4 and two batches of 2 round differently; measured 0.3e-2 at full resolution,
    # 5.3e-2 on head_0 -- the same level as the gradient-vs-oracle tests), losses within 2e-2
    vals = sorted(r0["grad_err"].values())
    assert vals[len(vals) // 2] < 5e-2 and vals[-1] < 1e-1, r0["grad_err"]
    assert torch.allclose(r0["losses_sharded_mean"], r0["losses_full"], rtol=2e-2, atol=2e-2), (r0["losses_sharded_mean"], r0["losses_full"])
    # graph replay == eager under data parallelism: the first iteration runs the same kernels on the same data
    assert r0["graph_vs_eager"][0] <= 2e-2 * float(r0["eager_losses"].abs().max()) + 2e-2, r0["graph_vs_eager"]
    assert r0["graph_vs_eager"][1] <= 5e-2 * float(r0["eager_losses"].abs().max()) + 2e-2, r0["graph_vs_eager"]
