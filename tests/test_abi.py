"""CPU-side checks of the boundary: the C-ABI library loads without a GPU and exports every symbol that
include/sgk.h declares (no compute calls), the Python surface mirrors the reference's networks API, and the
product package never imports the oracle."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    import supervised_gan_b200 as S
    return S


def test_header_symbols_are_exported(built):
    hdr = open(os.path.join(ROOT, "include", "sgk.h")).read()
    declared = set(re.findall(r"\b(sgk_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = built._lib.load()
    nm = subprocess.run(["nm", "-D", "--defined-only", built._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (sgk_[a-z0-9_]+)", nm))
    assert declared <= exported, declared - exported
    assert declared == set(built._lib.SIGNATURES), declared ^ set(built._lib.SIGNATURES)
    assert lib.sgk_version() == 1
    assert lib.sgk_adam_block_elems() == 4096


def test_plan_queries_need_no_gpu(built):
    import ctypes
    L = built._lib
    lib = L.load()
    d = L.SgkConvDesc(8, 128, 65, 65, 256, 66, 66, 4, 1, 2, 0, 0)
    assert lib.sgk_conv_packed_weight_elems(ctypes.byref(d), L.OP_FWD) == 256 * 128 * 16
    assert lib.sgk_conv_packed_weight_elems(ctypes.byref(d), L.OP_DGRAD) == 256 * 128 * 16
    assert lib.sgk_conv_wgrad_workspace_bytes(ctypes.byref(d)) > 0
    bad = L.SgkConvDesc(8, 128, 65, 65, 256, 60, 66, 4, 1, 2, 0, 0)   # wrong Hout
    assert lib.sgk_conv_packed_weight_elems(ctypes.byref(bad), L.OP_FWD) == 0
    t = L.SgkConvDesc(8, 256, 16, 16, 128, 32, 32, 4, 2, 1, 1, 0)     # ConvTranspose k4s2p1: 4 phases of 2x2 taps
    assert lib.sgk_conv_packed_weight_elems(ctypes.byref(t), L.OP_FWD) == 256 * 128 * 16


def test_python_surface_mirrors_reference(built):
    nw = built.networks
    import inspect
    sig = inspect.signature(nw.define_G)
    assert list(sig.parameters)[:6] == ["input_nc", "output_nc", "ngf", "which_model_netG", "norm", "use_dropout"]
    for name in ("n_layers_G", "use_residual", "use_fcn", "noise_nc", "add_gaussian_noise", "gaussian_sigma",
                 "n_layers_G_skip", "upsample_mode", "share_label_weights", "n_layers_CRN_block", "gpu_ids"):
        assert name in sig.parameters
    sigd = inspect.signature(nw.define_D)
    assert list(sigd.parameters) == ["input_nc", "ndf", "which_model_netD", "n_layers_D", "norm", "use_sigmoid",
                                     "scale_factor", "num_classes", "gpu_ids"]
    for name in ("GANLoss", "WeightedL1Loss", "print_network", "get_norm_layer", "weights_init", "FCGANGenerator",
                 "NLayerDiscriminator", "UnetGenerator", "UnetSkipConnectionBlock", "CascadedRefinementNetwork"):
        assert hasattr(nw, name)


def test_state_dict_keys_match_reference(built):
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    import torch
    ref = ref_loader.load()
    nw = built.networks
    cases = [
        (lambda m: m.define_G(2, 0, 32, "fcgan", "instance", False, n_layers_G=5, use_fcn=True, noise_nc=8, gpu_ids=[]), {}),
        (lambda m: m.define_G(2, 1, 8, "unet_256", "instance", True, gpu_ids=[]), {}),
        (lambda m: m.define_G(1, 2, 8, "unet_128", "instance", False, gpu_ids=[]), {}),
        (lambda m: m.define_G(2, 1, 16, "crn", "instance", False, n_layers_G=5, noise_nc=8, upsample_mode="bilinear",
                              n_layers_CRN_block=2, gpu_ids=[]), {}),
        (lambda m: m.define_G(2, 1, 16, "crn", "instance", False, n_layers_G=5, noise_nc=8, upsample_mode="convt",
                              share_label_weights=False, gpu_ids=[]), {}),
    ]
    for make, _ in cases:
        torch.manual_seed(5); a = make(ref)
        torch.manual_seed(5); b = make(nw)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert sa[k].shape == sb[k].shape and sa[k].dtype == sb[k].dtype
            assert torch.equal(sa[k], sb[k]), k      # same RNG consumption => identical initial weights
    for s in (1, 2, 4):
        torch.manual_seed(6); a = ref.define_D(3, 16, "n_layers", n_layers_D=4, norm="instance", use_sigmoid=True,
                                               scale_factor=ref_loader.sf(s), gpu_ids=[])
        torch.manual_seed(6); b = nw.define_D(3, 16, "n_layers", n_layers_D=4, norm="instance", use_sigmoid=True,
                                              scale_factor=s, gpu_ids=[])
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert torch.equal(sa[k], sb[k]), k


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "supervised-gan_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "/root/reference" not in src or f == "networks.py" or f.endswith(".py"), f


def test_cpu_input_fails_loudly(built):
    import torch
    D = built.networks.define_D(2, 8, "n_layers", n_layers_D=3, norm="instance", use_sigmoid=True, gpu_ids=[])
    with pytest.raises(RuntimeError, match="CUDA"):
        D(torch.zeros(1, 2, 32, 32))


def test_remaining_architectures_have_reference_state_dict_layout():
    """define_G / define_D of the architectures added for SURVEY 8f rank 4 expose the reference's state_dict keys and shapes
    (fixtures written by oracle/gen_golden_f4.py from the unmodified reference)."""
    import numpy as np
    import supervised_gan_b200 as S
    nw = S.networks
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    cases = {
        "f4_resnet6": lambda: nw.define_G(2, 1, 4, "resnet_6blocks", "instance", False, gpu_ids=[]),
        "f4_resnet9_res": lambda: nw.define_G(2, 2, 4, "resnet_9blocks", "instance", False, use_residual=True, gpu_ids=[]),
        "f4_autoencoder": lambda: nw.define_G(2, 1, 4, "autoencoder", "instance", False, n_layers_G=3, gpu_ids=[]),
        "f4_fcgan_star": lambda: nw.define_G(2, 0, 4, "fcgan_star", "instance", False, n_layers_G=5, use_fcn=True, noise_nc=8, gpu_ids=[]),
        "f4_dcgan_G": lambda: nw.define_G(3, 0, 8, "dcgan", "instance", False, noise_nc=8, gpu_ids=[]),
        "f4_dcgan_D": lambda: nw.define_D(3, 8, "dcgan", gpu_ids=[]),
        "f4_nlayersep_s2": lambda: nw.define_D(3, 4, "n_layers_sep", n_layers_D=3, norm="instance", use_sigmoid=True, scale_factor=2, gpu_ids=[]),
    }
    for name, make in cases.items():
        g = np.load(os.path.join(gdir, name + ".npz"))
        ref = {k[3:]: g[k].shape for k in g.files if k.startswith("sd.")}
        ours = {k: tuple(v.shape) for k, v in make().state_dict().items()}
        assert ours == ref, name
