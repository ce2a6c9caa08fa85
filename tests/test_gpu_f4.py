"""GPU parity of the remaining architectures (ResnetGenerator, AutoEncoder, FCGANGeneratorStar, DCGAN pair,
NLayerDiscriminatorSep, GANLossMultiClass, tensor2im) against fixtures produced by the UNMODIFIED reference
(oracle/gen_golden_f4.py), fp32 CUDA-core path, same tolerances as tests/test_gpu_nets.py."""
import numpy as np
import pytest
import torch

from tests.test_gpu_nets import check_module, sd_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import supervised_gan_b200 as S
    S.set_precision("fp32")
    return S


def zero_biases(g):
    """Conv biases that feed a norm: the reference's gradient is rounding noise of an exact zero (DESIGN.md deviations)."""
    out = []
    for k in g.files:
        if k.startswith("grad.") and k.endswith(".bias"):
            w = "grad." + k[5:-4] + "weight"
            if w in g.files and np.abs(g[k]).max() < 1e-3 * max(np.abs(g[w]).max(), 1e-3):
                out.append(k[5:])
    return out


def test_resnet_generators(S, golden):
    g = golden("f4_resnet6")
    G = S.networks.define_G(2, 1, 4, "resnet_6blocks", "instance", False, gpu_ids=[])
    check_module(G, g, ["x"], lambda n, x: n(x), zero_bias=zero_biases(g))
    g = golden("f4_resnet9_res")
    G = S.networks.define_G(2, 2, 4, "resnet_9blocks", "instance", False, use_residual=True, gpu_ids=[])
    check_module(G, g, ["x"], lambda n, x: n(x), zero_bias=zero_biases(g))


def test_autoencoder_and_fcgan_star(S, golden):
    g = golden("f4_autoencoder")
    G = S.networks.define_G(2, 1, 4, "autoencoder", "instance", False, n_layers_G=3, gpu_ids=[])
    check_module(G, g, ["x"], lambda n, x: n(x), zero_bias=zero_biases(g))
    g = golden("f4_fcgan_star")
    G = S.networks.define_G(2, 0, 4, "fcgan_star", "instance", False, n_layers_G=5, use_fcn=True, noise_nc=8, gpu_ids=[])
    check_module(G, g, ["z"], lambda n, z: n(z), zero_bias=zero_biases(g))


def test_dcgan_pair(S, golden):
    g = golden("f4_dcgan_G")
    G = S.networks.define_G(3, 0, 8, "dcgan", "instance", False, noise_nc=8, gpu_ids=[])
    check_module(G, g, ["z"], lambda n, z: n(z), zero_bias=zero_biases(g))
    g = golden("f4_dcgan_D")
    D = S.networks.define_D(3, 8, "dcgan", gpu_ids=[])
    check_module(D, g, ["x"], lambda n, x: n(x), zero_bias=zero_biases(g))


@pytest.mark.parametrize("s", [1, 2])
def test_nlayer_discriminator_sep(S, golden, s):
    g = golden("f4_nlayersep_s%d" % s)
    D = S.networks.define_D(3, 4, "n_layers_sep", n_layers_D=3, norm="instance", use_sigmoid=True, scale_factor=s, gpu_ids=[])
    check_module(D, g, ["x"], lambda n, x: n(x), zero_bias=zero_biases(g))


def test_multiclass_gan_loss_and_tensor2im(S, golden):
    g = golden("f4_ce_loss")
    crit = S.networks.GANLossMultiClass(num_classes=3)
    for t in (0, 2):
        x = torch.from_numpy(g["in.x"].copy()).cuda().requires_grad_(True)
        l = crit(x, t)
        l.backward()
        np.testing.assert_allclose(float(l), float(g["out.loss_%d" % t]), rtol=2e-6)
        np.testing.assert_allclose(x.grad.cpu().numpy(), g["out.grad_%d" % t], rtol=2e-5, atol=1e-8)
    # util.tensor2im (util/util.py:15-25): numpy restatement on the same values, 1 / 2 / 3 channels
    rng = np.random.RandomState(3)
    for C in (1, 2, 3):
        a = rng.uniform(-1, 1, size=(2, C, 9, 13)).astype(np.float32)
        ref = (a[0] + 1) / 2.0 * 255.0
        if C == 1:
            ref = ref.repeat(3, 0)
        elif C == 2:
            ref = np.concatenate((ref, np.zeros([1, 9, 13], dtype=ref.dtype)), axis=0)
        ref = np.transpose(ref, (1, 2, 0)).astype(np.uint8)
        got = S.ops.tensor2im(torch.from_numpy(a).cuda())
        assert got.dtype == np.uint8 and np.array_equal(got, ref), C


def test_reflection_pad_matches_torch(S):
    """op-level: forward bit-exact, backward vs torch's own ReflectionPad2d autograd (the checker) on odd shapes."""
    import torch.nn.functional as F
    for (N, C, H, W, p) in [(2, 4, 9, 7, 3), (1, 3, 5, 6, 1), (2, 8, 4, 4, 3)]:
        x = torch.randn(N, C, H, W, device="cuda")
        xr = x.clone().requires_grad_(True)
        yr = F.pad(xr, (p, p, p, p), mode="reflect")
        gy = torch.randn_like(yr)
        yr.backward(gy)
        xn = x.permute(0, 2, 3, 1).contiguous().requires_grad_(True)
        y = S.ops.reflection_pad(xn, p)
        y.backward(gy.permute(0, 2, 3, 1).contiguous())
        assert torch.equal(y.permute(0, 3, 1, 2), yr)
        assert (xn.grad.permute(0, 3, 1, 2) - xr.grad).abs().max() <= 1e-5


def test_inference_sampler_visuals(S, tmp_path):
    """test.py:40-49 for the fcgan model: save a trained-for-one-step model, reload it with isTrain=False, model.test(), and
    get_current_visuals() -> util.tensor2im images (numpy restatement of util/util.py:15-25 as the checker)."""
    from tests.test_gpu_step import make_opt
    from supervised_gan_b200.fcgan_model import FCGANModel
    torch.manual_seed(3)
    kw = dict(which_channel="r_g", ngf=8, ndf=8, noiseSize=2, fineSize=128, scale_factor=[1, 2], lambda_D=[0.6, 0.4],
              n_layers_D=[3, 3], checkpoints_dir=str(tmp_path), name="s", pool_size=0)
    m = FCGANModel(); m.initialize(make_opt(**kw))
    m.set_input({"A": torch.rand(1, 3, 128, 128) * 2 - 1, "A_paths": ["x"]})
    m.optimize_parameters()
    vis = m.get_current_visuals()
    assert list(vis) == ["real_label", "real_image", "fake_label", "fake_image"]
    m.save("latest")
    t = FCGANModel(); t.initialize(make_opt(isTrain=False, **kw))
    t.test()
    vis = t.get_current_visuals()
    assert list(vis) == ["fake_label", "fake_image"]
    fake = t.fake.detach().cpu().numpy()
    for k, c in (("fake_label", 0), ("fake_image", 1)):
        ref = ((fake[0, c:c + 1] + 1) / 2.0 * 255.0).repeat(3, 0)
        ref = np.transpose(ref, (1, 2, 0)).astype(np.uint8)
        assert vis[k].dtype == np.uint8 and vis[k].shape == (128, 128, 3) and np.array_equal(vis[k], ref)
