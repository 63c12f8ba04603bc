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
        self.grad_hook = None   # optional callable(list of params) run before each optimizer.step (data-parallel all-reduce)

    def generator_losses(self, real_A, real_B, masks=None):
        """trainer.py:447-512: returns (loss_G, dict of the individual terms, fake_A, fake_B)."""
        cat = (lambda t: torch.cat([t, masks], dim=1)) if masks is not None else (lambda t: t)
        real_A_input, real_B_input = cat(real_A), cat(real_B)
        fake_B, fake_A, id_A, id_B = self._translate_and_identity(real_A_input, real_B_input)
        loss_id = (l1_loss(id_A, real_A) + l1_loss(id_B, real_B)) / 2
        loss_GAN = (mse_gan_loss(self.D_B(fake_B), True) + mse_gan_loss(self.D_A(fake_A), True)) / 2
        rec_A, rec_B = self.G_B2A(cat(fake_B)), self.G_A2B(cat(fake_A))
        loss_cycle = (l1_loss(rec_A, real_A) + l1_loss(rec_B, real_B)) / 2
        loss_grad_cycle = (self.criterion_gradient(rec_A, real_A) + self.criterion_gradient(rec_B, real_B)) / 2
        loss_grad_id = (self.criterion_gradient(id_A, real_A) + self.criterion_gradient(id_B, real_B)) / 2
        loss_ssim = 1 - ((self.criterion_ssim(rec_A, real_A) + self.criterion_ssim(rec_B, real_B)) / 2)
        loss_att = self.criterion_contrast_attention(fake_B, real_B, real_A)
        loss_region = self.criterion_contrast_region(fake_B, real_B, real_A)
        loss_edge = self.criterion_contrast_edge(fake_B, real_B, real_A)
        loss_G = (loss_GAN + self.lambda_cyc * loss_cycle + self.lambda_id * loss_id + 5.0 * loss_grad_cycle + 2.5 * loss_grad_id
                  + 2.0 * loss_ssim + 2.0 * loss_att + 1.5 * loss_region + 1.0 * loss_edge)
        terms = dict(GAN=loss_GAN, cycle=loss_cycle, id=loss_id, grad_cycle=loss_grad_cycle, grad_id=loss_grad_id, ssim=loss_ssim,
                     contrast_attention=loss_att, contrast_region=loss_region, contrast_edge=loss_edge)
        return loss_G, terms, fake_A, fake_B

    def _translate_and_identity(self, real_A_input, real_B_input):
        """trainer.py:464-467: fake_B, fake_A = G_A2B(A), G_B2A(B); id_A, id_B = G_B2A(A), G_A2B(B).  Each generator sees its two
        inputs as ONE batch of 2B samples (InstanceNorm is per sample and the kernels are batch-invariant, so every sample's
        output and gradient are what the two separate calls give; half the launches, twice the work per launch)."""
        B = real_A_input.shape[0]
        ab = self.G_A2B(torch.cat([real_A_input, real_B_input], dim=0))
        ba = self.G_B2A(torch.cat([real_B_input, real_A_input], dim=0))
        return ab[:B], ba[:B], ba[B:], ab[B:]

    def _disc_loss(self, D, real, fake):
        """trainer.py:518 / :523: (MSE(D(real), valid) + MSE(D(fake.detach()), fake)) / 2, both images in one batch of 2B."""
        B = real.shape[0]
        out = D(torch.cat([real, fake.detach()], dim=0))
        return (mse_gan_loss(out[:B], True) + mse_gan_loss(out[B:], False)) / 2

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
