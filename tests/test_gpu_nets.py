"""GPU parity, network level: our drop-in networks (define_G / define_D, libsgk kernels) against
 (a) the committed golden fixtures produced by the UNMODIFIED reference modules, and
 (b) the oracle (oracle/nets.py, fp32 and fp64 on CPU) at the real config-1 widths.
Weights are loaded through load_state_dict with the reference's keys, proving state_dict compatibility.
fp32 CUDA-core path tolerances: outputs 2e-5 abs (tanh/sigmoid range), parameter gradients 3e-4 relative
to the tensor's max (fp32 reference itself is ~1e-4..1e-3 from fp64 on these ill-conditioned sums)."""
import numpy as np
import pytest
import torch

from oracle import nets as ON
from tests.conftest import grad_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import supervised_gan_b200 as S
    S.set_precision("fp32")
    return S


def sd_of(g, prefix="sd"):
    return {k[len(prefix) + 1:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith(prefix + ".")}


def relerr(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    return np.abs(np.asarray(got, dtype=np.float64) - ref).max() / max(np.abs(ref).max(), 1e-7)


ZERO_GRAD_BIAS_ATOL = 1e-6   # conv biases that feed a norm: exact 0 here, ~1e-9 noise in the reference


def check_module(net, g, inputs, call, out_tol=3e-6, grad_tol=1e-3, zero_bias=()):
    net.load_state_dict(sd_of(g))
    net.cuda()
    ins = {k: torch.from_numpy(g["in." + k].copy()).cuda().requires_grad_(True) for k in inputs}
    y = call(net, **ins)
    assert y.shape == g["out.y"].shape
    assert np.abs(y.detach().cpu().numpy() - g["out.y"]).max() <= out_tol
    (y * torch.from_numpy(g["in.proj"]).cuda()).sum().backward()
    params = dict(net.named_parameters())
    for k in g.files:
        if k.startswith("grad."):
            name = k[5:]
            if name.startswith("gauss_filter"):
                continue  # never optimised (fcgan_model.py:100-109); we do not compute this wasted gradient
            got = params[name].grad.cpu().numpy()
            if name in zero_bias:
                assert np.abs(got).max() <= ZERO_GRAD_BIAS_ATOL, name   # reference: rounding noise of a zero gradient
            else:
                grad_close(got, g[k], name, q_tol=grad_tol)
        if k.startswith("gin."):
            grad_close(ins[k[4:]].grad.cpu().numpy(), g[k], k, q_tol=grad_tol)
        if k.startswith("sd_after.") and "running" in k:
            np.testing.assert_allclose(net.state_dict()[k[9:]].cpu().numpy(), g[k], rtol=2e-5, atol=1e-6)
        if k.startswith("sd_after.") and "tracked" in k:
            assert int(net.state_dict()[k[9:]]) == int(g[k])


def test_fcgan_generator_golden(S, golden):
    nw = S.networks
    G = nw.define_G(2, 0, 4, "fcgan", "instance", False, n_layers_G=5, use_fcn=True, noise_nc=8, gpu_ids=[])
    check_module(G, golden("fcgan_G_fcn"), ["z"], lambda n, z: n(z), zero_bias=["model.%d.bias" % i for i in (3, 6, 9, 12)])
    G = nw.define_G(2, 0, 4, "deconv", "instance", False, n_layers_G=4, use_fcn=False, noise_nc=8, gpu_ids=[])
    check_module(G, golden("fcgan_G_nofcn"), ["z"], lambda n, z: n(z), zero_bias=["model.%d.bias" % i for i in (3, 6, 9)])


@pytest.mark.parametrize("s,nl,sig", [(1, 3, True), (2, 3, True), (4, 3, True), (1, 4, False), (2, 2, False)])
def test_nlayer_discriminator_golden(S, golden, s, nl, sig):
    D = S.networks.define_D(2, 4, "n_layers", n_layers_D=nl, norm="instance", use_sigmoid=sig, scale_factor=s, gpu_ids=[])
    zero = ["model.%d.bias" % (2 + 3 * i) for i in range(nl)]
    check_module(D, golden("nlayerD_s%d_n%d_%s" % (s, nl, "sig" if sig else "lin")), ["x"], lambda n, x: n(x), zero_bias=zero)


def test_unet_golden(S, golden):
    U = S.networks.define_G(2, 1, 2, "unet_128", "instance", False, gpu_ids=[])
    zb = [k for k in U.state_dict() if k.endswith(".bias") and k not in ("model.0.bias", "model.3.bias")]
    # the innermost down-conv (no norm after it) keeps a real bias gradient
    inner = "model.1" + ".model.3" * 5 + ".model.1.bias"
    zb.remove(inner)
    check_module(U, golden("unet128"), ["x"], lambda n, x: n(x), zero_bias=zb)
    U = S.networks.define_G(1, 2, 2, "unet_256", "instance", False, gpu_ids=[])
    zb = [k for k in U.state_dict() if k.endswith(".bias") and k not in ("model.0.bias", "model.3.bias")]
    zb.remove("model.1" + ".model.3" * 6 + ".model.1.bias")
    check_module(U, golden("unet256"), ["x"], lambda n, x: n(x), zero_bias=zb)


def test_crn_golden(S, golden):
    for mode, nb in (("bilinear", 2), ("convt", 1)):
        C = S.networks.define_G(2, 1, 8, "crn", "instance", False, n_layers_G=5, noise_nc=8, upsample_mode=mode,
                                n_layers_CRN_block=nb, gpu_ids=[])
        last = "blockh0.1.model.%d.bias" % (3 * (nb - 1) + 1)
        zb = [k for k in C.state_dict() if k.endswith(".bias") and k != last]
        check_module(C, golden("crn_%s_b%d" % (mode, nb)), ["label", "noise"], lambda n, label, noise: n(label, noise),
                     zero_bias=zb)


def test_config1_widths_vs_oracle_fp32_and_fp64(S):
    """Real config-1 widths (ngf 32 / ndf 32, 3 scales) at 128x128, B=2, loss_G = sum lambda*BCE(D_s(G(z)), 1) at fixed
    weights: outputs/loss within 4x the fp32 oracle's own error to fp64; every G gradient within the robust bound of
    conftest.grad_close against fp64 (the same bound the fp32 oracle is held to)."""
    gen = torch.Generator().manual_seed(11)
    sdG = ON.init_fcgan_generator(gen, 8, 2, 32, 5)
    sdDs = [ON.init_nlayer_discriminator(gen, 2, 32, 3, s) for s in (1, 2, 4)]
    z = torch.randn(2, 8, 2, 2, generator=gen)
    lam = (0.5, 0.4, 0.1)

    def oracle(dtype):
        sg = {k: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sdG.items()}
        sds = [{k: v.clone().to(dtype) for k, v in sd.items()} for sd in sdDs]
        for d in [sg] + sds:
            for k, v in d.items():
                if v.is_floating_point() and "running" not in k:
                    v.requires_grad_(True)
        fake = ON.fcgan_generator(sg, z.to(dtype), 5, True)
        loss = sum(l * ON.gan_loss(ON.nlayer_discriminator(sd, fake, 3, s, True), True) for l, sd, s in zip(lam, sds, (1, 2, 4)))
        loss.backward()
        return fake.detach(), float(loss), {k: v.grad for k, v in sg.items() if v.requires_grad}

    f64, l64, g64 = oracle(torch.float64)
    f32, l32, g32 = oracle(torch.float32)

    nw = S.networks
    G = nw.define_G(2, 0, 32, "fcgan", "instance", False, n_layers_G=5, use_fcn=True, noise_nc=8, gpu_ids=[])
    G.load_state_dict(sdG); G.cuda()
    Ds = []
    for s, sd in zip((1, 2, 4), sdDs):
        D = nw.define_D(2, 32, "n_layers", n_layers_D=3, norm="instance", use_sigmoid=True, scale_factor=s, gpu_ids=[])
        D.load_state_dict(sd); D.cuda(); Ds.append(D)
    crit = nw.GANLoss(use_lsgan=False)
    fake = G(z.cuda())
    loss = 0
    for l, D in zip(lam, Ds):
        loss = loss + crit(D(fake), True) * l
    loss.backward()
    assert relerr(fake.detach().cpu().numpy(), f64.numpy()) <= 4 * relerr(f32.numpy(), f64.numpy()) + 2e-6
    assert abs(float(loss) - l64) <= 4 * abs(l32 - l64) + 2e-6
    zero = {"model.%d.bias" % i for i in (3, 6, 9, 12)}
    for k, p in G.named_parameters():
        if k in zero:
            assert float(p.grad.abs().max()) == 0.0
            continue
        grad_close(p.grad.cpu().numpy(), g64[k].numpy(), k)          # ours vs fp64 truth
        grad_close(g32[k].numpy(), g64[k].numpy(), "oracle fp32 " + k)  # the reference's arithmetic meets the same bar


def test_state_dict_round_trip_and_api_surface(S):
    nw = S.networks
    D = nw.define_D(3, 8, "basic", norm="instance", use_sigmoid=False, scale_factor=2, gpu_ids=[])
    assert D.gauss_filter is not None and hasattr(D, "model") and D.gpu_ids == []
    assert tuple(D.state_dict()["gauss_filter.0.weight"].shape) == (3, 3, 5, 5)
    with pytest.raises(NotImplementedError):
        nw.define_G(2, 1, 8, "nope", "instance")
    with pytest.raises(NotImplementedError):
        nw.define_D(2, 8, "nope")
    with pytest.raises(NotImplementedError):
        nw.get_norm_layer("layer")
    # arbitrary callable as the final activation, as the reference's forward(x, activation=...) allows
    G = nw.define_G(2, 0, 4, "fcgan", "instance", False, n_layers_G=2, use_fcn=True, noise_nc=4, gpu_ids=[]).cuda()
    z = torch.randn(1, 4, 2, 2, device="cuda")
    a = G(z, activation=torch.nn.Tanh())
    G.model[1].num_batches_tracked.zero_()
    b = G(z, activation=lambda t: t)
    assert torch.allclose(torch.tanh(b), a, atol=1e-6)
