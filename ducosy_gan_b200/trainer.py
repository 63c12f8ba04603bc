"""One CycleGAN optimisation step, statement for statement what reference modules/trainer.py:447-525 does with the
models, criteria and optimisers built at trainer.py:320-362 -- on the kernels of libducosy_sm100.so.

The only torch tensor ops left are the ones the reference's own loop body performs between modules: ``torch.cat`` of
the mask channels (trainer.py:450-452,476-478) and the scalar arithmetic of the loss mix (trainer.py:493-512).
One scheduling difference, with identical per-sample results: where the reference calls the same network twice on
independent inputs (translation + identity, trainer.py:464-467; real + fake, trainer.py:518,523) the two inputs go through
as one batch.
"""
from __future__ import annotations

import torch

from .losses import (SSIM, ContrastAttentionLoss, ContrastEdgeLoss, ContrastRegionLoss, GradientLoss, l1_loss, mse_gan_loss)
from .modules.model import Discriminator, Generator, weights_init_normal
from .optim import Adam


class CycleGANStep:
    """Holds G_A2B, G_B2A, D_A, D_B, the nine criteria and the three Adam optimisers (trainer.py:328-362)."""

    def __init__(self, input_channels=1, num_residual_blocks=9, use_cbam=True, lr=2e-4, lambda_cyc=10.0, lambda_id=5.0,
                 device="cuda", seed=None, capturable=False):
        if seed is not None:
            torch.manual_seed(seed)
        dev = torch.device(device)
        self.G_A2B = Generator(input_channels, num_residual_blocks, use_cbam).to(dev)
        self.G_B2A = Generator(input_channels, num_residual_blocks, use_cbam).to(dev)
        self.D_A = Discriminator(1).to(dev)
        self.D_B = Discriminator(1).to(dev)
        for m in (self.G_A2B, self.G_B2A, self.D_A, self.D_B):   # trainer.py:409-417
            m.apply(weights_init_normal)
        self.criterion_gradient = GradientLoss()
        self.criterion_ssim = SSIM(data_range=1.0, size_average=True, channel=1)
        self.criterion_contrast_attention = ContrastAttentionLoss(sigma=0.15, min_weight=1.0, max_weight=3.0, blur_kernel=7)
        self.criterion_contrast_region = ContrastRegionLoss(threshold=0.15, weight=1.5)
        self.criterion_contrast_edge = ContrastEdgeLoss()
        self.optimizer_G = Adam(list(self.G_A2B.parameters()) + list(self.G_B2A.parameters()), lr=lr, betas=(0.5, 0.999),
                                capturable=capturable)
        self.optimizer_D_A = Adam(self.D_A.parameters(), lr=lr, betas=(0.5, 0.999), capturable=capturable)
        self.optimizer_D_B = Adam(self.D_B.parameters(), lr=lr, betas=(0.5, 0.999), capturable=capturable)
        self.lambda_cyc, self.lambda_id = lambda_cyc, lambda_id
        self._streams = {}
        self.grad_hook = None   # optional callable(list of params) run before each optimizer.step (data-parallel all-reduce)

    def _side_streams(self, device):
        """Two side streams, one per generator (DUCOSY_TRAIN_STREAMS=1 runs everything on the caller's stream).  G_A2B and
        G_B2A are independent in the translation / identity pass and again in the cycle pass, and a generator pass alternates
        tensor-core convolutions with bandwidth-bound normalisation / padding kernels, so two passes side by side fill each
        other's gaps -- the more so the smaller the per-rank batch.  Autograd runs every backward node on the stream of its
        forward, so the backward is two-stream as well; a CUDA-graph capture keeps the fork / join structure."""
        import os
        if os.environ.get("DUCOSY_TRAIN_STREAMS", "2") == "1" or device.type != "cuda":
            return None
        key = device.index if device.index is not None else torch.cuda.current_device()
        if key not in self._streams:
            with torch.cuda.device(key):
                self._streams[key] = (torch.cuda.Stream(), torch.cuda.Stream())
        return self._streams[key]

    def _batch_global_losses(self, fake_B, real_B, real_A):
        """The two criteria that take statistics over the WHOLE batch tensor (trainer.py:117-127,163-181); returns
        (loss_region, loss_edge, weight of the pair in the expression that is differentiated).  The data-parallel step
        overrides this (gathered batch, weight = world)."""
        return (self.criterion_contrast_region(fake_B, real_B, real_A), self.criterion_contrast_edge(fake_B, real_B, real_A), 1.0)

    def generator_losses(self, real_A, real_B, masks=None):
        """trainer.py:447-512: returns (loss_G, dict of the individual terms, fake_A, fake_B).

        Scheduling differences, with identical per-sample results: where the reference calls the same network twice on
        independent inputs (translation + identity, trainer.py:464-467) the two inputs go through as one batch of 2B (the
        kernels are batch invariant), and the two generators run on two side streams (``_side_streams``)."""
        cat = (lambda t: torch.cat([t, masks], dim=1)) if masks is not None else (lambda t: t)
        real_A_input, real_B_input = cat(real_A), cat(real_B)
        B = real_A.shape[0]
        streams = self._side_streams(real_A.device)
        if streams is None:
            ab = self.G_A2B(torch.cat([real_A_input, real_B_input], dim=0))
            ba = self.G_B2A(torch.cat([real_B_input, real_A_input], dim=0))
            fake_B, id_B, fake_A, id_A = ab[:B], ab[B:], ba[:B], ba[B:]
            rec_A, rec_B = self.G_B2A(cat(fake_B)), self.G_A2B(cat(fake_A))
            join = lambda: None
        else:
            cur = torch.cuda.current_stream()
            s1, s2 = streams                       # s1: every G_A2B pass, s2: every G_B2A pass
            s1.wait_stream(cur)
            s2.wait_stream(cur)
            with torch.cuda.stream(s1):
                ab = self.G_A2B(torch.cat([real_A_input, real_B_input], dim=0))
                e1 = s1.record_event()
            with torch.cuda.stream(s2):
                ba = self.G_B2A(torch.cat([real_B_input, real_A_input], dim=0))
                e2 = s2.record_event()
            fake_B, id_B, fake_A, id_A = ab[:B], ab[B:], ba[:B], ba[B:]
            ab.record_stream(cur), ab.record_stream(s2), ba.record_stream(cur), ba.record_stream(s1)
            s1.wait_event(e2)                      # the cycle pass of each generator reads the other one's translation
            s2.wait_event(e1)
            with torch.cuda.stream(s2):
                rec_A = self.G_B2A(cat(fake_B))
            with torch.cuda.stream(s1):
                rec_B = self.G_A2B(cat(fake_A))
            rec_A.record_stream(cur), rec_B.record_stream(cur)
            cur.wait_event(e1)                     # identity / adversarial / contrast terms overlap the cycle passes
            cur.wait_event(e2)

            def join():
                cur.wait_stream(s1)
                cur.wait_stream(s2)
        loss_id = (l1_loss(id_A, real_A) + l1_loss(id_B, real_B)) / 2
        loss_GAN = (mse_gan_loss(self.D_B(fake_B), True) + mse_gan_loss(self.D_A(fake_A), True)) / 2
        loss_grad_id = (self.criterion_gradient(id_A, real_A) + self.criterion_gradient(id_B, real_B)) / 2
        loss_att = self.criterion_contrast_attention(fake_B, real_B, real_A)
        loss_region, loss_edge, w_global = self._batch_global_losses(fake_B, real_B, real_A)
        join()
        loss_cycle = (l1_loss(rec_A, real_A) + l1_loss(rec_B, real_B)) / 2
        loss_grad_cycle = (self.criterion_gradient(rec_A, real_A) + self.criterion_gradient(rec_B, real_B)) / 2
        loss_ssim = 1 - ((self.criterion_ssim(rec_A, real_A) + self.criterion_ssim(rec_B, real_B)) / 2)
        loss_G = (loss_GAN + self.lambda_cyc * loss_cycle + self.lambda_id * loss_id + 5.0 * loss_grad_cycle + 2.5 * loss_grad_id
                  + 2.0 * loss_ssim + 2.0 * loss_att + w_global * (1.5 * loss_region + 1.0 * loss_edge))
        terms = dict(GAN=loss_GAN, cycle=loss_cycle, id=loss_id, grad_cycle=loss_grad_cycle, grad_id=loss_grad_id, ssim=loss_ssim,
                     contrast_attention=loss_att, contrast_region=loss_region, contrast_edge=loss_edge)
        return loss_G, terms, fake_A, fake_B

    def _disc_loss(self, D, real, fake):
        """trainer.py:518 / :523: (MSE(D(real), valid) + MSE(D(fake.detach()), fake)) / 2, both images in one batch of 2B."""
        B = real.shape[0]
        out = D(torch.cat([real, fake.detach()], dim=0))
        return (mse_gan_loss(out[:B], True) + mse_gan_loss(out[B:], False)) / 2

    # ------------------------------------------------------------------ validation (trainer.py:187-294)
    @torch.no_grad()
    def validation_loss(self, batches):
        """``validate_and_save_images`` part 1 (trainer.py:205-253): mean over the batches of
        ``loss_GAN + lambda_cyc * loss_cycle + lambda_id * loss_id`` with both generators in eval mode and no autograd.
        ``batches``: iterable of dicts with "A", "B" (and optionally "masks") as the reference's dataloader yields them.
        The three losses of a batch stay on the device; one host read at the end."""
        was_training = self.G_A2B.training, self.G_B2A.training
        self.G_A2B.eval(), self.G_B2A.eval()
        dev = next(self.G_A2B.parameters()).device
        total, n = None, 0
        try:
            for batch in batches:
                real_A, real_B = batch["A"].to(dev), batch["B"].to(dev)
                masks = batch["masks"].to(dev) if "masks" in batch else None
                cat = (lambda t: torch.cat([t, masks], dim=1)) if masks is not None else (lambda t: t)
                B = real_A.shape[0]
                ab = self.G_A2B(torch.cat([cat(real_A), cat(real_B)], dim=0))     # fake_B | id_B   (trainer.py:228-241)
                ba = self.G_B2A(torch.cat([cat(real_B), cat(real_A)], dim=0))     # fake_A | id_A
                fake_B, id_B, fake_A, id_A = ab[:B], ab[B:], ba[:B], ba[B:]
                rec_A, rec_B = self.G_B2A(cat(fake_B)), self.G_A2B(cat(fake_A))
                loss_id = (l1_loss(id_A, real_A) + l1_loss(id_B, real_B)) / 2
                loss_GAN = (mse_gan_loss(self.D_B(fake_B), True) + mse_gan_loss(self.D_A(fake_A), True)) / 2
                loss_cycle = (l1_loss(rec_A, real_A) + l1_loss(rec_B, real_B)) / 2
                loss_G = loss_GAN + self.lambda_cyc * loss_cycle + self.lambda_id * loss_id
                total = loss_G if total is None else total + loss_G
                n += 1
        finally:
            self.G_A2B.train(was_training[0]), self.G_B2A.train(was_training[1])
        return float(total) / max(n, 1) if n else 0.0

    @torch.no_grad()
    def validation_image_grid(self, batch, hu_min, hu_max, window_center, window_width):
        """``validate_and_save_images`` part 2 (trainer.py:257-280): real_A | fake_B | real_B, each display-windowed
        (preprocess.py:58-65), concatenated along the width -> [B,1,H,3W] fp32 in [0,1] (what ``save_image`` receives)."""
        from . import ops
        dev = next(self.G_A2B.parameters()).device
        was_training = self.G_A2B.training
        self.G_A2B.eval()
        try:
            real_A, real_B = batch["A"].to(dev), batch["B"].to(dev)
            x = torch.cat([real_A, batch["masks"].to(dev)], dim=1) if "masks" in batch else real_A
            fake_B = self.G_A2B(x)
        finally:
            self.G_A2B.train(was_training)
        win = lambda t: ops.apply_windowing(t, hu_min, hu_max, window_center, window_width)
        return torch.cat((win(real_A), win(fake_B), win(real_B)), -1)

    def _sync(self, opt):
        if self.grad_hook is not None:
            self.grad_hook([p for g in opt.param_groups for p in g["params"]])

    def step(self, real_A, real_B, masks=None):
        """One iteration of the loop body (trainer.py:447-525).  Returns a dict of detached loss tensors (no host sync)."""
        self.optimizer_G.zero_grad()
        loss_G, terms, fake_A, fake_B = self.generator_losses(real_A, real_B, masks)
        loss_G.backward()
        self._sync(self.optimizer_G)
        self.optimizer_G.step()

        self.optimizer_D_A.zero_grad()
        loss_D_A = self._disc_loss(self.D_A, real_A, fake_A)
        loss_D_A.backward()
        self._sync(self.optimizer_D_A)
        self.optimizer_D_A.step()

        self.optimizer_D_B.zero_grad()
        loss_D_B = self._disc_loss(self.D_B, real_B, fake_B)
        loss_D_B.backward()
        self._sync(self.optimizer_D_B)
        self.optimizer_D_B.step()
        out = {k: v.detach() for k, v in terms.items()}
        out.update(G=loss_G.detach(), D_A=loss_D_A.detach(), D_B=loss_D_B.detach())
        return out
