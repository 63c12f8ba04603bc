"""CycleGAN step (reference modules/trainer.py:447-525) on the CUDA path: the fused Adam against
torch.optim.Adam, and the step-1 loss terms against the oracle restatement evaluated on the CPU with the same weights."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ducosy_oracle as orc  # noqa: E402


def test_fused_adam_matches_torch_adam():
    from ducosy_gan_b200.optim import Adam
    torch.manual_seed(0)
    shapes = [(256, 256, 3, 3), (64,), (1, 2, 7, 7), (1000003,)]
    mine = [torch.randn(s, device="cuda").requires_grad_(True) for s in shapes]
    ref = [p.detach().clone().requires_grad_(True) for p in mine]
    o1, o2 = Adam(mine, lr=2e-4, betas=(0.5, 0.999)), torch.optim.Adam(ref, lr=2e-4, betas=(0.5, 0.999))
    sched = torch.optim.lr_scheduler.LambdaLR(o1, lambda e: 1.0 - 0.1 * e)       # trainer.py:364-366 drives lr this way
    sched2 = torch.optim.lr_scheduler.LambdaLR(o2, lambda e: 1.0 - 0.1 * e)
    for it in range(4):
        for a, b in zip(mine, ref):
            g = torch.randn_like(a) * (10.0 ** (it - 2))
            a.grad, b.grad = g.clone(), g.clone()
        v0 = mine[0]._version
        o1.step(); o2.step(); sched.step(); sched2.step()
        assert mine[0]._version > v0      # the packed 16-bit weight caches of the modules are keyed by _version
        for a, b in zip(mine, ref):
            assert (a - b).abs().max().item() <= 1e-6 * b.abs().max().item() + 1e-7


@pytest.mark.parametrize("capturable", [False, True])
def test_fused_adam_state_dict_round_trip(capturable):
    """checkpoint / resume as reference modules/trainer.py:394-396,586-588 does it: an optimiser restored from state_dict()
    continues exactly like the uninterrupted one (including the device-side step count of the capturable mode), and a
    state_dict written by torch.optim.Adam loads into the fused one."""
    from ducosy_gan_b200.optim import Adam
    torch.manual_seed(3)
    w0 = torch.randn(300, 70, device="cuda")
    grads = [torch.randn_like(w0) for _ in range(5)]

    def run(opt, p, gs):
        for g in gs:
            p.grad = g.clone()
            opt.step()

    a = w0.clone().requires_grad_(True)
    oa = Adam([a], lr=1e-3, betas=(0.5, 0.999), capturable=capturable)
    run(oa, a, grads)                                             # uninterrupted
    b = w0.clone().requires_grad_(True)
    ob = Adam([b], lr=1e-3, betas=(0.5, 0.999), capturable=capturable)
    run(ob, b, grads[:2])
    import copy
    ckpt = copy.deepcopy(ob.state_dict())
    c = b.detach().clone().requires_grad_(True)
    oc = Adam([c], lr=1e-3, betas=(0.5, 0.999), capturable=capturable)
    oc.load_state_dict(ckpt)
    run(oc, c, grads[2:])
    assert torch.equal(a, c)
    t = w0.clone().requires_grad_(True)                           # a checkpoint written by the reference's optimiser class
    ot = torch.optim.Adam([t], lr=1e-3, betas=(0.5, 0.999))
    run(ot, t, grads[:2])
    d = t.detach().clone().requires_grad_(True)
    od = Adam([d], lr=1e-3, betas=(0.5, 0.999), capturable=capturable)
    import copy
    od.load_state_dict(copy.deepcopy(ot.state_dict()))        # state_dict() hands out the live buffers, not copies
    run(od, d, grads[2:])
    run(ot, t, grads[2:])
    assert (d - t).abs().max().item() <= 1e-6 * t.abs().max().item() + 1e-7


def test_fused_adam_invalidates_packed_weight_caches():
    """the kernel writes parameters through raw pointers: a module's next forward must see the new weights"""
    from ducosy_gan_b200.modules.model import Discriminator, weights_init_normal
    from ducosy_gan_b200.optim import Adam
    torch.manual_seed(1)
    D = Discriminator(1).cuda().apply(weights_init_normal)
    x = torch.rand(1, 1, 256, 256, device="cuda") * 2 - 1
    with torch.no_grad():
        y0 = D(x).clone()
    opt = Adam(D.parameters(), lr=1e-2, betas=(0.5, 0.999))
    D(x).square().mean().backward()
    opt.step()
    with torch.no_grad():
        y1 = D(x)
    assert (y1 - y0).abs().max().item() > 1e-3


def _oracle_step_losses(sdG_A, sdG_B, sdD_A, sdD_B, real_A, real_B, masks, blocks, cbam):
    """step-1 loss terms from the oracle's restatement of the loop body (oracle.cyclegan_generator_loss / _discriminator_loss)"""
    loss_G, t, fake_A, fake_B = orc.cyclegan_generator_loss(sdG_A, sdG_B, sdD_A, sdD_B, real_A, real_B, masks, blocks, cbam)
    t = dict(t, G=loss_G, D_A=orc.cyclegan_discriminator_loss(sdD_A, real_A, fake_A),
             D_B=orc.cyclegan_discriminator_loss(sdD_B, real_B, fake_B))
    return {k: float(v) for k, v in t.items()}


@pytest.mark.parametrize("cfg", [(2, 1, True, 1), (1, 2, False, 2)])
def test_train_step_losses_match_oracle_and_parameters_move(cfg):
    """Step-1 loss terms (computed before any update) against the oracle on the same weights and batch: 3 % relative
    (16-bit activations through up to three chained networks; |.|-type losses are first-order in that noise) -- D_A / D_B
    are evaluated after optimizer_G.step() in the reference loop, but on fakes made before it, so the oracle's values with
    the initial discriminator weights are exact references for them too.  Then every parameter must have moved and two
    further steps must stay finite."""
    from ducosy_gan_b200.trainer import CycleGANStep
    Cin, blocks, cbam, B = cfg
    H, W = 256, 512            # the PatchGAN kernels need multiples of 256
    step = CycleGANStep(Cin, blocks, cbam, seed=11)
    g = torch.Generator().manual_seed(3)
    smooth = lambda t: torch.nn.functional.avg_pool2d(t, 5, 1, 2) * 2.0
    real_A = smooth(torch.rand(B, 1, H, W, generator=g) * 2 - 1).clamp(-1, 1)
    real_B = smooth(torch.rand(B, 1, H, W, generator=g) * 2 - 1).clamp(-1, 1)
    masks = (torch.rand(B, Cin - 1, H, W, generator=g) < 0.1).float() if Cin > 1 else None
    cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    sds = [cpu(m) for m in (step.G_A2B, step.G_B2A, step.D_A, step.D_B)]
    before = [p.detach().clone() for m in (step.G_A2B, step.G_B2A, step.D_A, step.D_B) for p in m.parameters()]
    with torch.no_grad():
        ref = _oracle_step_losses(*sds, real_A, real_B, masks, blocks, cbam)
    dev = lambda t: None if t is None else t.cuda()
    got = {k: v.item() for k, v in step.step(dev(real_A), dev(real_B), dev(masks)).items()}
    bad = {k: (got[k], ref[k]) for k in ref if abs(got[k] - ref[k]) > 3e-2 * abs(ref[k]) + 1e-4}
    assert not bad, bad
    after = [p for m in (step.G_A2B, step.G_B2A, step.D_A, step.D_B) for p in m.parameters()]
    names = [n for m in (step.G_A2B, step.G_B2A, step.D_A, step.D_B) for n, _ in m.named_parameters()]
    live = [(n, a, b) for n, a, b in zip(names, after, before)]
    still = [n for n, a, b in live if torch.equal(a, b) and not (n.endswith("bias") and a.numel() > 1)]
    assert not still, still            # dead biases (in front of an InstanceNorm) legitimately keep a zero gradient
    for _ in range(2):
        out = step.step(dev(real_A), dev(real_B), dev(masks))
    assert all(torch.isfinite(v).item() for v in out.values()), out


def test_train_loss_trajectory_matches_oracle():
    """BASELINE north_star: "training losses after N steps".  Four optimisation steps of the CUDA path against four steps of
    the fp32 oracle (autograd through the restated networks / losses, torch.optim.Adam) from the same initial weights on
    the same batch.  Adam's first updates are +-lr per weight, so the 16-bit gradient noise of the CUDA path only flips
    updates of near-zero gradients and the trajectories stay together: every logged loss within 5 % at every step."""
    from ducosy_gan_b200.trainer import CycleGANStep
    Cin, blocks, cbam, B, H, W, N = 1, 1, True, 1, 256, 512, 4
    step = CycleGANStep(Cin, blocks, cbam, seed=21)
    g = torch.Generator().manual_seed(8)
    smooth = lambda t: torch.nn.functional.avg_pool2d(t, 5, 1, 2) * 2.0
    real_A = smooth(torch.rand(B, 1, H, W, generator=g) * 2 - 1).clamp(-1, 1)
    real_B = smooth(torch.rand(B, 1, H, W, generator=g) * 2 - 1).clamp(-1, 1)
    cpu = lambda m: {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    sdGA, sdGB, sdDA, sdDB = (cpu(m) for m in (step.G_A2B, step.G_B2A, step.D_A, step.D_B))
    oG = torch.optim.Adam(list(sdGA.values()) + list(sdGB.values()), lr=2e-4, betas=(0.5, 0.999))
    oDA = torch.optim.Adam(list(sdDA.values()), lr=2e-4, betas=(0.5, 0.999))
    oDB = torch.optim.Adam(list(sdDB.values()), lr=2e-4, betas=(0.5, 0.999))
    ref_hist, got_hist = [], []
    for _ in range(N):
        # ---- oracle step (reference modules/trainer.py:462-524 restated in oracle.cyclegan_step)
        ref = orc.cyclegan_step((sdGA, sdGB, sdDA, sdDB), (oG, oDA, oDB), real_A, real_B, None, blocks, cbam)
        ref_hist.append({k: ref[k] for k in ("G", "D_A", "D_B", "GAN", "id")})
        # ---- CUDA step
        out = step.step(real_A.cuda(), real_B.cuda())
        got_hist.append({k: out[k].item() for k in ref_hist[-1]})
    bad = [(i, k, got_hist[i][k], ref_hist[i][k]) for i in range(N) for k in ref_hist[i]
           if abs(got_hist[i][k] - ref_hist[i][k]) > 5e-2 * abs(ref_hist[i][k]) + 1e-4]
    assert not bad, (bad, got_hist, ref_hist)
    assert ref_hist[-1]["G"] < ref_hist[0]["G"] and got_hist[-1]["G"] < got_hist[0]["G"]      # both actually train


def test_graphed_step_matches_eager_step():
    """The CUDA-graph replay of the whole optimisation step against the eager step: same weights, same batches, five steps
    (three of them are the warm-up the capture needs).  Kernels are deterministic, so the losses agree to float rounding of
    Adam's bias correction (host double vs device double pow): 1e-4 relative."""
    from ducosy_gan_b200.data_parallel import DataParallelCycleGANStep, GraphedCycleGANStep
    g = torch.Generator().manual_seed(4)
    batches = [((torch.rand(1, 1, 256, 512, generator=g) * 2 - 1).cuda(), (torch.rand(1, 1, 256, 512, generator=g) * 2 - 1).cuda())
               for _ in range(3)]
    eager = DataParallelCycleGANStep(1, 1, True, seed=31, capturable=False)
    ref = [eager.step(*batches[0]) for _ in range(3)]            # what the three warm-up steps of the capture do
    ref += [eager.step(*batches[1]), eager.step(*batches[2])]
    ref = [{k: v.item() for k, v in r.items()} for r in ref]
    step = DataParallelCycleGANStep(1, 1, True, seed=31, capturable=True)
    graphed = GraphedCycleGANStep(step, *batches[0], warmup=3)
    got = []
    for a, b in batches[1:]:
        # the capture itself does not execute; the first replay is step 4
        out = graphed(a, b)
        got.append({k: v.item() for k, v in out.items()})
    for r, o in zip(ref[3:], got):
        bad = {k: (o[k], r[k]) for k in r if abs(o[k] - r[k]) > 1e-4 * abs(r[k]) + 1e-6}
        assert not bad, bad
    # weights really moved and the version counters tell the modules so
    with torch.no_grad():
        y = step.D_A(batches[0][0])
        y2 = eager.D_A(batches[0][0])
    assert (y - y2).abs().max().item() < 1e-3 * y2.abs().max().item() + 1e-5


def _config4_case():
    """BASELINE configs[3] network at batch 1: soft-tissue generators (input_channels = 3: slice + bone / mediastinum masks,
    9 CBAM blocks), PatchGAN discriminators, 512x512."""
    from ducosy_gan_b200.trainer import CycleGANStep
    step = CycleGANStep(3, 9, True, seed=1234)
    g = torch.Generator().manual_seed(2)
    smooth = lambda t: torch.nn.functional.avg_pool2d(t, 5, 1, 2) * 2.0
    real_A = smooth(torch.rand(1, 1, 512, 512, generator=g) * 2 - 1).clamp(-1, 1)
    real_B = smooth(torch.rand(1, 1, 512, 512, generator=g) * 2 - 1).clamp(-1, 1)
    masks = (torch.rand(1, 2, 512, 512, generator=g) < 0.1).float()
    return step, real_A, real_B, masks


def _oracle_step1(sds, real_A, real_B, masks):
    """losses + per-tensor gradients of step 1 from oracle.cyclegan_generator_loss / cyclegan_discriminator_loss under autograd"""
    for sd in sds.values():
        for v in sd.values():
            v.grad = None
    loss_G, t, fake_A, fake_B = orc.cyclegan_generator_loss(sds["G_A2B"], sds["G_B2A"], sds["D_A"], sds["D_B"], real_A, real_B, masks, 9, True)
    loss_G.backward()
    grads = {n: {k: v.grad.clone() for k, v in sds[n].items()} for n in ("G_A2B", "G_B2A")}
    for n in ("D_A", "D_B"):
        for v in sds[n].values():
            v.grad = None
    l_DA = orc.cyclegan_discriminator_loss(sds["D_A"], real_A, fake_A)
    l_DA.backward()
    l_DB = orc.cyclegan_discriminator_loss(sds["D_B"], real_B, fake_B)
    l_DB.backward()
    for n in ("D_A", "D_B"):
        grads[n] = {k: v.grad.clone() for k, v in sds[n].items()}
    losses = {k: float(v.detach()) for k, v in t.items()}
    losses.update(G=float(loss_G.detach()), D_A=float(l_DA.detach()), D_B=float(l_DB.detach()))
    return losses, grads


def test_config4_network_step1_losses_and_gradients_vs_oracle(monkeypatch):
    """Step 1 of the config-4 network (Cin 3, 9 CBAM blocks, 512x512, batch 1) against oracle.cyclegan_generator_loss /
    cyclegan_discriminator_loss under torch autograd (fp32, CPU): all 12 logged terms, and the gradient of EVERY parameter
    tensor of the four networks (relative L2 per tensor, cosine per network).

    Two references.  (a) the plain fp32 oracle.  The network is piecewise linear -- 23 ReLU gates per generator pass, the CBAM
    max-pools, the sign maps of the seven |.|-type loss terms -- and those gates are decided by the forward values: a forward
    with 10-bit-mantissa operands (fp16 here, TF32 in the reference's own GPU runs) moves a small fraction f of the pixels
    across a gate at every layer, and flipping a fraction f of equal-magnitude entries is a relative L2 change of 2*sqrt(f).
    tools/grad_noise_experiment.py (profiles/r02_grad_noise_experiment.txt) separates the two roundings on a plain torch
    model: rounding the FORWARD alone gives 5-6 % per tensor at the stem, rounding the gradient maps alone < 0.1 %.
    (b) the same oracle with the generator forward rounded where the kernels store 16 bit (straight-through gradient), which
    brings the gates of the reference close to the kernels' and leaves mostly the backward arithmetic.
    Floors measured on the B200 (fp16 operands): profiles/r02_config4_step1.json."""
    import json
    import os
    import test_gpu_gen_backward as TG
    step, real_A, real_B, masks = _config4_case()
    nets = dict(G_A2B=step.G_A2B, G_B2A=step.G_B2A, D_A=step.D_A, D_B=step.D_B)
    sds = {n: {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()} for n, m in nets.items()}
    ref, ref_grads = _oracle_step1(sds, real_A, real_B, masks)
    q = TG._ste_round(torch.float16)
    monkeypatch.setattr(orc, "generator_forward", lambda sd, x, nb=9, cbam=True: TG._torch_generator(sd, x, nb, cbam, q))
    _, aware_grads = _oracle_step1(sds, real_A, real_B, masks)
    monkeypatch.undo()
    # ---- CUDA path: the same statements through the product modules
    dev = lambda x: x.cuda()
    for m in nets.values():
        m.zero_grad(set_to_none=True)
    cg, ct, cfA, cfB = step.generator_losses(dev(real_A), dev(real_B), dev(masks))
    cg.backward()
    got_grads = {n: {k: p.grad.detach().cpu().clone() for k, p in nets[n].named_parameters()} for n in ("G_A2B", "G_B2A")}
    for n in ("D_A", "D_B"):
        nets[n].zero_grad(set_to_none=True)         # trainer.py:517,522 discard what loss_G.backward() left there
    c_DA = step._disc_loss(step.D_A, dev(real_A), cfA)
    c_DA.backward()
    c_DB = step._disc_loss(step.D_B, dev(real_B), cfB)
    c_DB.backward()
    for n in ("D_A", "D_B"):
        got_grads[n] = {k: p.grad.detach().cpu().clone() for k, p in nets[n].named_parameters()}
    got = {k: float(v) for k, v in ct.items()}
    got.update(G=float(cg), D_A=float(c_DA), D_B=float(c_DB))
    bad = {k: (got[k], ref[k]) for k in ref if abs(got[k] - ref[k]) > 1e-3 * abs(ref[k]) + 1e-5}
    print("config-4 step-1 losses (cuda, oracle):", {k: (round(got[k], 5), round(ref[k], 5)) for k in ref})
    assert not bad, bad                              # measured: every term within 4e-4 relative

    def compare(reference):
        table, cos = {}, {}
        for n in got_grads:
            gs, rs = [], []
            for k, g in got_grads[n].items():
                r = reference[n][k]
                dead = k.endswith("bias") and g.abs().max().item() == 0.0 and r.norm().item() < 1e-3 * max(v.norm().item() for v in reference[n].values())
                if dead:                             # bias in front of a non-affine InstanceNorm: exact zero here, rounding noise there
                    continue
                table[f"{n}.{k}"] = ((g - r).norm() / (r.norm() + 1e-30)).item()
                gs.append(g.reshape(-1))
                rs.append(r.reshape(-1))
            cos[n] = torch.nn.functional.cosine_similarity(torch.cat(gs), torch.cat(rs), dim=0).item()
        return table, cos

    table, cos = compare(ref_grads)
    table_aware, cos_aware = compare(aware_grads)
    small = lambda k: "cbam" in k          # 98- and 4096-element attention tensors: sums of strongly cancelling terms
    summary = lambda t: {"conv_weights_max": max(v for k, v in t.items() if not small(k)), "attention_tensors_max": max(v for k, v in t.items() if small(k)),
                         "median": sorted(t.values())[len(t) // 2]}
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump({"losses_cuda_vs_oracle": {k: [got[k], ref[k]] for k in ref},
               "vs_fp32_oracle": {"summary": summary(table), "cosine_per_network": cos, "grad_rel_l2": table},
               "vs_rounding_aware_oracle": {"summary": summary(table_aware), "cosine_per_network": cos_aware, "grad_rel_l2": table_aware}},
              open(os.path.join(out, "config4_step1.json"), "w"), indent=1)
    print("vs fp32 oracle:          ", summary(table), {k: round(v, 5) for k, v in cos.items()})
    print("vs rounding-aware oracle:", summary(table_aware), {k: round(v, 5) for k, v in cos_aware.items()})
    # (a) fp32 oracle: sign-flip floor of the |.|-type losses (measured 0.10 on every conv weight, 0.49 worst attention tensor)
    s = summary(table)
    assert s["conv_weights_max"] < 0.15 and s["attention_tensors_max"] < 0.75 and s["median"] < 0.12, s
    assert table["G_B2A.model.28.weight"] < 0.02 and table["G_A2B.model.28.weight"] < 0.02     # last layer: no chain behind it
    assert max(v for k, v in table.items() if k.startswith("D_")) < 0.08
    assert min(cos.values()) > 0.99, cos
    # (b) rounding-aware oracle: tighter (most sign noise removed)
    s = summary(table_aware)
    assert s["conv_weights_max"] < 0.15 and s["median"] < 0.12, s
    assert min(cos_aware.values()) > 0.99, cos_aware


def test_adam_skips_a_step_with_non_finite_gradients():
    """GradScaler-style guard of the fused Adam: a step whose gradients hold Inf/NaN changes neither the parameters nor the
    moments nor the step count; the next clean step is exactly the step an unpoisoned optimiser would have taken."""
    from ducosy_gan_b200.optim import Adam
    torch.manual_seed(0)
    shapes = [(64, 32, 3, 3), (5,), (1,), (70001,)]
    a = [torch.randn(s, device="cuda").requires_grad_(True) for s in shapes]
    b = [p.detach().clone().requires_grad_(True) for p in a]
    oa, ob = Adam(a, lr=1e-3, betas=(0.5, 0.999)), Adam(b, lr=1e-3, betas=(0.5, 0.999))
    grads = [[torch.randn(s, device="cuda") for s in shapes] for _ in range(3)]
    for k in range(3):
        if k == 1:                                   # poisoned step for `a` only
            for p, g in zip(a, grads[k]):
                p.grad = g.clone()
            a[3].grad[12345] = float("nan")
            a[0].grad[3, 2, 1, 0] = float("inf")
            before = [p.detach().clone() for p in a]
            oa.step()
            assert all(torch.equal(x, y) for x, y in zip(before, a))
            assert oa.skipped_steps() == 1
            continue
        for p, q, g in zip(a, b, grads[k]):
            p.grad, q.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert oa.state_dict()["state"][0]["step"] == 2 and ob.skipped_steps() == 0


def test_training_survives_a_constant_channel():
    """A dead conv output channel (all-zero weights) is a constant map: InstanceNorm's backward scales its gradient by
    1/sqrt(0 + 1e-5) = 316, the regime the 16-bit gradient maps are most exposed in.  Whatever happens there, the master
    weights must stay finite (the update is skipped if a gradient overflowed)."""
    from ducosy_gan_b200.trainer import CycleGANStep
    step = CycleGANStep(1, 2, True, seed=5)
    with torch.no_grad():
        for G in (step.G_A2B, step.G_B2A):
            G.model[1].weight[7].zero_()                       # stem channel 7 dead
            G.model[4].weight[11].zero_()                      # down-conv channel 11 dead
            G.model[10].block[1].weight[3].zero_()             # first res-block conv, channel 3 dead
            G.model[10].block[5].weight[3].zero_()
    g = torch.Generator().manual_seed(1)
    real_A = (torch.rand(1, 1, 256, 512, generator=g) * 2 - 1).cuda()
    real_B = (torch.rand(1, 1, 256, 512, generator=g) * 2 - 1).cuda()
    for _ in range(3):
        out = step.step(real_A, real_B)
    assert all(torch.isfinite(v).item() for v in out.values()), out
    for m in (step.G_A2B, step.G_B2A, step.D_A, step.D_B):
        assert all(torch.isfinite(p).all().item() for p in m.parameters())
    print("skipped steps:", step.optimizer_G.skipped_steps(), step.optimizer_D_A.skipped_steps(), step.optimizer_D_B.skipped_steps())


def test_validation_loss_and_image_grid_match_oracle():
    """validate_and_save_images (reference modules/trainer.py:187-294): the validation loss
    GAN + 10 cycle + 5 id averaged over two batches (eval mode, no autograd) against the oracle's restatement of the same
    terms, and the display-windowed real_A | fake_B | real_B grid against oracle.apply_windowing; training mode is restored."""
    from ducosy_gan_b200.trainer import CycleGANStep
    Cin, blocks, cbam, B, H, W = 2, 1, True, 1, 256, 512
    step = CycleGANStep(Cin, blocks, cbam, seed=5)
    g = torch.Generator().manual_seed(9)
    smooth = lambda t: torch.nn.functional.avg_pool2d(t, 5, 1, 2) * 2.0
    batches = []
    for _ in range(2):
        batches.append({"A": smooth(torch.rand(B, 1, H, W, generator=g) * 2 - 1).clamp(-1, 1),
                        "B": smooth(torch.rand(B, 1, H, W, generator=g) * 2 - 1).clamp(-1, 1),
                        "masks": (torch.rand(B, Cin - 1, H, W, generator=g) < 0.1).float()})
    cpu = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    sds = [cpu(m) for m in (step.G_A2B, step.G_B2A, step.D_A, step.D_B)]
    want = 0.0
    with torch.no_grad():
        for b in batches:
            _, t, _, fake_B = orc.cyclegan_generator_loss(*sds, b["A"], b["B"], b["masks"], blocks, cbam)
            want += float(t["GAN"] + 10.0 * t["cycle"] + 5.0 * t["id"])        # trainer.py:249
    want /= len(batches)
    step.G_A2B.train(), step.G_B2A.train()
    got = step.validation_loss(batches)
    assert step.G_A2B.training and step.G_B2A.training
    print(f"validation loss: cuda {got:.5f}  oracle {want:.5f}")
    assert abs(got - want) <= 3e-2 * abs(want)
    grid = step.validation_image_grid(batches[-1], -150, 250, 40, 400).cpu()
    assert grid.shape == (B, 1, H, 3 * W)
    win = lambda x: torch.from_numpy(orc.apply_windowing(x, -150, 250, 40, 400))
    ref = torch.cat((win(batches[-1]["A"]), win(fake_B), win(batches[-1]["B"])), -1)
    assert torch.equal(grid[..., :W], ref[..., :W]) and torch.equal(grid[..., 2 * W:], ref[..., 2 * W:])   # real images: same fp32 steps
    assert (grid - ref).abs().max().item() < 1.5e-2        # fake_B: the generator's stated fp16 tolerance (window scale 400/400)
