"""Checkpoint interop (reference: models/base_model.py:44-61).

CPU (`-m "not gpu"`, needs /root/reference, skipped elsewhere): a '<epoch>_net_<label>.pth' written by OUR save_network is
loaded by the UNMODIFIED reference's BaseModel.load_network into the reference's own module, and the other way round, with
bit-identical tensors.
GPU: train, save() (networks + Adam moments / step counts), resume in a fresh model with continue_train -- the resumed run
must continue bit-identically, which fails if the moments or the bias-correction step were lost."""
import argparse
import os
import types

import pytest
import torch

from oracle import nets as ON


def test_pth_interop_with_reference(tmp_path):
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    import supervised_gan_b200 as S
    from supervised_gan_b200.fcgan_model import FCGANModel
    ref = ref_loader.load()
    RefModel = ref_loader.load_model_class("fcgan")
    nw = S.networks
    mk_g = lambda m: m.define_G(2, 0, 8, "fcgan", "instance", False, n_layers_G=5, use_fcn=True, noise_nc=8, gpu_ids=[])
    mk_d = lambda m, s: m.define_D(2, 8, "n_layers", n_layers_D=3, norm="instance", use_sigmoid=True, scale_factor=s, gpu_ids=[])
    ours_self = types.SimpleNamespace(save_dir=str(tmp_path / "ours"), device=torch.device("cpu"))
    ref_self = RefModel()
    ref_self.save_dir = str(tmp_path / "ref")
    os.makedirs(ref_self.save_dir)
    for label, make_ours, make_ref in (("G", lambda: mk_g(nw), lambda: mk_g(ref)),
                                       ("D_1", lambda: mk_d(nw, 2), lambda: mk_d(ref, ref_loader.sf(2)))):
        # ours -> reference
        torch.manual_seed(1); a = make_ours()
        torch.manual_seed(2); b = make_ref()
        FCGANModel.save_network(ours_self, a, label, "7")
        ref_self.save_dir = ours_self.save_dir
        ref_self.load_network(b, label, "7")
        for (k, u), (k2, v) in zip(a.state_dict().items(), b.state_dict().items()):
            assert k == k2 and torch.equal(u, v), k
        # reference -> ours
        torch.manual_seed(3); b = make_ref()
        torch.manual_seed(4); a = make_ours()
        ref_self.save_dir = str(tmp_path / "ref")
        ref_self.save_network(b, label, "9", gpu_ids=[])
        ours_self.save_dir = ref_self.save_dir
        FCGANModel.load_network(ours_self, a, label, "9")
        ours_self.save_dir = str(tmp_path / "ours")
        for (k, u), (k2, v) in zip(a.state_dict().items(), b.state_dict().items()):
            assert k == k2 and torch.equal(u, v), k


def _opt(ckpt, **kw):
    d = dict(isTrain=True, gpu_ids=[0], checkpoints_dir=ckpt, name="ck", pretrained_model_dir="",
             which_channel="rg", batchSize=2, output_nc=2, input_nc=2, fineSize=64, noise_nc=8, noiseSize=1, ngf=8,
             which_model_netG="fcgan", norm="instance", no_dropout=True, n_layers_G=5, use_residual=False,
             add_gaussian_noise=False, gaussian_sigma=0.1, upsample_mode="convt", n_layers_CRN_block=1,
             no_share_label_block_weights=False, no_lsgan=True, scale_factor=[1, 2, 4], lambda_D=[0.5, 0.4, 0.1],
             n_layers_D=[3, 3, 3], ndf=8, which_model_netD="n_layers", continue_train=False, which_epoch="latest",
             pool_size=0, lr=2e-4, beta1=0.5, which_direction="A", n_update_D=1, n_update_G=1, no_logD_trick=False,
             niter_decay=100)
    d.update(kw)
    return argparse.Namespace(**d)


@pytest.mark.gpu
def test_resume_with_optimizer_state_is_bit_identical(tmp_path):
    import supervised_gan_b200 as S
    from supervised_gan_b200.fcgan_model import FCGANModel
    S.set_precision("fp32")
    gen = torch.Generator().manual_seed(8)
    reals = [(torch.rand(2, 2, 64, 64, generator=gen) * 2 - 1).cuda() for _ in range(5)]
    noises = [torch.randn(2, 8, 1, 1, generator=gen).cuda() for _ in range(5)]

    def steps(m, ts):
        out = []
        for t in ts:
            m._draw_noise = lambda t=t: noises[t]
            m.input.copy_(reals[t])
            m.optimize_parameters()
            out.append([float(m.loss_G), float(m.loss_D_real), float(m.loss_D_fake)])
        return out

    torch.manual_seed(9)
    a = FCGANModel(); a.initialize(_opt(str(tmp_path)))
    steps(a, [0, 1, 2])
    a.save("3")
    files = sorted(os.listdir(os.path.join(str(tmp_path), "ck")))
    assert files == ["3_net_D_0.pth", "3_net_D_1.pth", "3_net_D_2.pth", "3_net_G.pth", "3_optim_D.pth", "3_optim_G.pth"], files
    # the network files are what the reference writes: a CPU state_dict under the reference's keys
    sdG = torch.load(os.path.join(str(tmp_path), "ck", "3_net_G.pth"))
    assert all(not v.is_cuda for v in sdG.values())
    tail_a = steps(a, [3, 4])

    b = FCGANModel(); b.initialize(_opt(str(tmp_path), continue_train=True, which_epoch="3"))
    assert b.optimizer_G.state_dict()["state"], "Adam moments were not restored"
    tail_b = steps(b, [3, 4])
    assert tail_a == tail_b, (tail_a, tail_b)
    assert b.optimizer_G.step_count() == 5 and b.optimizer_D.step_count() == 5
    for (k, u), (_, v) in zip(a.netG.state_dict().items(), b.netG.state_dict().items()):
        assert torch.equal(u, v), k
    for da, db in zip(a.netD, b.netD):
        for (k, u), (_, v) in zip(da.state_dict().items(), db.state_dict().items()):
            assert torch.equal(u, v), k

    # the saved generator, evaluated by the oracle restatement of the reference's FCGANGenerator, is the network we trained
    d = FCGANModel(); d.initialize(_opt(str(tmp_path), continue_train=True, which_epoch="3"))
    y_ref = ON.fcgan_generator({k: v.clone() for k, v in sdG.items()}, noises[0].cpu(), 5, use_fcn=False, update_running=False)
    assert (y_ref - d.netG(noises[0]).detach().cpu()).abs().max() <= 2e-5

    # without the optimiser files the resumed run restarts Adam cold and diverges immediately
    os.remove(os.path.join(str(tmp_path), "ck", "3_optim_G.pth"))
    os.remove(os.path.join(str(tmp_path), "ck", "3_optim_D.pth"))
    c = FCGANModel(); c.initialize(_opt(str(tmp_path), continue_train=True, which_epoch="3"))
    tail_c = steps(c, [3, 4])
    # (the D losses of the first resumed step are taken before any optimiser step; everything after differs)
    assert tail_c[0][1:] == tail_a[0][1:] and tail_c[0][0] != tail_a[0][0] and tail_c[1] != tail_a[1]
