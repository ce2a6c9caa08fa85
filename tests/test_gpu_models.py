"""GPU parity of the conditional-GAN and DSGAN two-stage step drivers (our CGANModel / TwoStageCycleModel on the libsgk
kernels) against golden fixtures produced by the UNMODIFIED reference model classes (oracle/gen_golden.step_fixtures).
Step-0 losses are tight (they do not depend on optimiser dynamics beyond one D update); later steps are loose for the
Adam sign-descent reason explained in DESIGN.md section 4."""
import argparse

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import supervised_gan_b200 as S
    S.set_precision("fp32")
    return S


def sd_of(g, prefix):
    return {k[len(prefix) + 1:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith(prefix + ".")}


def base_opt(**kw):
    d = dict(isTrain=True, gpu_ids=[0], checkpoints_dir="/tmp/sgk_ckpt", name="t", pretrained_model_dir="", batchSize=1,
             norm="instance", add_gaussian_noise=False, gaussian_sigma=0.1, continue_train=False, which_epoch="latest",
             pool_size=0, lr=2e-4, beta1=0.5, no_logD_trick=False, niter_decay=100, weights=None, no_cgan=False,
             dataset_mode="single", which_direction="AtoB", lambda_A=10.0, transform_1to2="none")
    d.update(kw)
    return argparse.Namespace(**d)


def test_cgan_step_golden(S, golden):
    from supervised_gan_b200.cgan_model import CGANModel
    g = golden("cgan_step")
    opt = base_opt(which_channel="rg_b", fineSize=128, noise_nc=8, noiseSize=8, ngf=2, ndf=4, which_model_netG="unet_128",
                   which_model_netD="n_layers", no_dropout=True, n_layers_G=5, use_residual=False, upsample_mode="convt",
                   n_layers_CRN_block=1, no_share_label_block_weights=False, n_layers_G_skip=-1, no_lsgan=True,
                   scale_factor=[1, 2], n_layers_D=[3, 2], lambda_D=[0.6, 0.4], weights=[2.0, 3.0], n_update_D=1, n_update_G=1,
                   input_nc=2, output_nc=1)
    m = CGANModel(); m.initialize(opt)
    m.netG.load_state_dict(sd_of(g, "sdG"))
    for i, d in enumerate(m.netD):
        d.load_state_dict(sd_of(g, "sdD%d" % i))
    S.ops.bump_weights_epoch()
    for t in range(int(g["meta.steps"])):
        m.input_A = torch.from_numpy(g["in.real_A%d" % t]).cuda()
        m.input_B = torch.from_numpy(g["in.real_B%d" % t]).cuda()
        m.optimize_parameters()
        e = m.get_current_errors()
        got = [e["G_GAN"], e["G_L1"], e["D_real"], e["D_fake"]]
        np.testing.assert_allclose(got, g["out.loss%d" % t], rtol=1e-4 if t == 0 else 5e-3, atol=2e-6)
        if t == 0:
            assert np.abs(m.fake_B.detach().cpu().numpy() - g["out.fake_B0"]).max() <= 5e-6
    lr, steps = 2e-4, int(g["meta.steps"])
    for k, v in m.netG.state_dict().items():
        d = np.abs(v.cpu().numpy() - g["sdG_after." + k])
        assert d.max() <= 2.2 * lr * steps, k


def test_twostage_cycle_step_golden(S, golden):
    from supervised_gan_b200.twostage_cycle_model import TwoStageCycleModel
    g = golden("twostage_step")
    opt = base_opt(which_channel="rg_b", fineSize=128, input_nc=2, output_nc=1, noise_nc1=8, noiseSize1=1, noise_nc2=8,
                   noiseSize2=2, ngf1=4, ngf2=8, nff2=2, ndf1=4, ndf2=4, which_model_netG1="fcgan", which_model_netG2="crn",
                   which_model_netF2="unet_128", which_model_netD1="n_layers", which_model_netD2="n_layers",
                   which_model_netD="n_layers", n_layers_G1=4, n_layers_G2=5, n_layers_F2=5, no_dropout1=True,
                   no_dropout2=True, use_residual2=False, upsample_mode1="convt", upsample_mode2="bilinear",
                   n_layers_CRN_block1=1, n_layers_CRN_block2=2, no_share_label_block_weights1=False,
                   no_share_label_block_weights2=False, transform_1to2="bilinear_2", scale_factor1=[1, 2], lambda_D1=[0.5, 0.4],
                   n_layers_D1=[3, 2], scale_factor2=[1, 2], lambda_D2=[0.6, 0.4], n_layers_D2=[3, 3], no_lsgan1=True,
                   no_lsgan2=True, use_multi_class_GAN=False, use_fixed_noise1=False, sequential_train=False, lr1=2e-4, lr2=2e-4,
                   n_update_D1=1, n_update_D2=1, n_update_G=1, detach_G1_from_G2_x=False, detach_G1_from_G2_y=False,
                   GAN_losses_D2=["real_fake", "fake_fake"], GAN_losses_G2=["real_fake", "fake_fake"], lambda_B=10.0,
                   lambda_A_cycle=5.0, lambda_fake_cycle=1.0)
    m = TwoStageCycleModel(); m.initialize(opt)
    for lab, net in (("G1", m.netG1), ("G2", m.netG2), ("F2", m.netF2)):
        net.load_state_dict(sd_of(g, "sd" + lab))
    for i, d in enumerate(m.netD1):
        d.load_state_dict(sd_of(g, "sdD1_%d" % i))
    for i, d in enumerate(m.netD2):
        d.load_state_dict(sd_of(g, "sdD2_%d" % i))
    S.ops.bump_weights_epoch()
    for t in range(int(g["meta.steps"])):
        m.input_A = torch.from_numpy(g["in.real_A%d" % t]).cuda()
        m.input_B = torch.from_numpy(g["in.real_B%d" % t]).cuda()
        n1, n2 = torch.from_numpy(g["in.noise1_%d" % t]).cuda(), torch.from_numpy(g["in.noise2_%d" % t]).cuda()
        m._draw_noises = lambda: (n1, n2)
        m.optimize_parameters()
        e = m.get_current_errors()
        got = [float(m.loss_G), e["G1_GAN"], e["G2_GAN"], e["G2_L1"], e["F2_CE"], e["G2_real_cycle"], e["G2_fake_cycle"],
               e["D1_real"], e["D1_fake"], e["D2_real"], e["D2_fake"]]
        np.testing.assert_allclose(got, g["out.loss%d" % t], rtol=2e-4 if t == 0 else 1e-2, atol=5e-6)
        if t == 0:
            assert np.abs(m.fake_A.detach().cpu().numpy() - g["out.fake_A0"]).max() <= 5e-6
            assert np.abs(m.fake_B_from_fake_A.detach().cpu().numpy() - g["out.fake_B_from_fake_A0"]).max() <= 2e-5
            assert np.abs(m.recon_fake_A.detach().cpu().numpy() - g["out.recon_fake_A0"]).max() <= 2e-5
    # three-group optimiser honours per-group learning rates and the decay protocol
    m.update_learning_rate()
    assert abs(m.optimizer_G.param_groups[0]["lr"] - (2e-4 - 2e-4 / 100)) < 1e-12
