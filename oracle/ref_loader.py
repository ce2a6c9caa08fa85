"""Loader for the UNMODIFIED reference (phymhan/supervised-gan) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py and by the `-m "not gpu"`
tests that re-validate the oracle restatement against the real reference when the
reference tree is present (it only exists in the build container, never on the
GPU box).  Nothing under supervised-gan_b200/ may import this file.

Two non-invasive shims are required to run the 2017-era code under Python 3 /
torch 2.x:
  1. networks.py:127-128, 808-811 rely on Python-2 integer division of
     `scale_factor / 2`; we hand the reference an `int` subclass whose true
     division floors.
  2. util/util.py:9 imports skimage (absent); an empty stub module is registered.
"""
import os
import sys
import types
import argparse

_VENDORED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
# /root/reference in the build container; on the GPU box the byte-for-byte copy made by oracle/vendor_ref.py
REF_ROOT = os.environ.get("SGK_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/models/networks.py") else _VENDORED)


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "networks.py"))


class Py2Int(int):
    """int whose `/` is Python-2 floor division (networks.py:127, 808)."""

    def __truediv__(self, o):
        return Py2Int(int(self) // int(o))

    def __mul__(self, o):
        return Py2Int(int(self) * int(o))

    __rmul__ = __mul__

    def __add__(self, o):
        return Py2Int(int(self) + int(o))

    __radd__ = __add__


def sf(s):
    return Py2Int(s) if s > 1 else s


_loaded = {}


def load():
    """Returns the reference's `models.networks` module."""
    if "networks" in _loaded:
        return _loaded["networks"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    for name in ("skimage", "skimage.measure"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from models import networks  # noqa: the reference's module
    _loaded["networks"] = networks
    return networks


def load_model_class(which):
    """Reference step drivers: 'fcgan' | 'cgan' | 'twostage_cycle'."""
    load()
    if which == "fcgan":
        from models.fcgan_model import FCGANModel
        return FCGANModel
    if which == "cgan":
        from models.cgan_model import CGANModel
        return CGANModel
    if which == "twostage_cycle":
        from models.twostage_cycle_model import TwoStageCycleModel
        return TwoStageCycleModel
    raise ValueError(which)


def fcgan_opt(**kw):
    """Namespace with the fields FCGANModel.initialize reads (fcgan_model.py:32-116)."""
    d = dict(isTrain=True, gpu_ids=[], checkpoints_dir="/tmp/sgk_ckpt", name="oracle",
             pretrained_model_dir="", which_channel="rg", batchSize=1, output_nc=2, input_nc=2,
             fineSize=512, noise_nc=8, noiseSize=8, ngf=32, which_model_netG="fcgan",
             norm="instance", no_dropout=True, n_layers_G=5, use_residual=False,
             add_gaussian_noise=False, gaussian_sigma=0.1, upsample_mode="convt",
             n_layers_CRN_block=1, no_share_label_block_weights=False, no_lsgan=True,
             scale_factor=[1, Py2Int(2), Py2Int(4)], lambda_D=[0.5, 0.4, 0.1],
             n_layers_D=[3, 3, 3], ndf=32, which_model_netD="n_layers", continue_train=False,
             which_epoch="latest", pool_size=50, lr=2e-4, beta1=0.5, which_direction="A",
             n_update_D=1, n_update_G=1, no_logD_trick=False, niter_decay=100)
    d.update(kw)
    d["scale_factor"] = [sf(int(s)) for s in d["scale_factor"]]
    return argparse.Namespace(**d)


def cgan_opt(**kw):
    """Fields CGANModel.initialize reads (cgan_model.py:18-115) on top of the fcgan ones."""
    d = dict(which_channel="rg_b", which_model_netG="unet_128", ngf=64, ndf=64, scale_factor=[1, 1], lambda_D=[0.5, 0.5],
             n_layers_D=[3, 4], transform_1to2="none", n_layers_G_skip=-1, no_cgan=False, weights=None,
             dataset_mode="single", lambda_A=10.0, which_direction="AtoB", noiseSize=8)
    d.update(kw)
    return fcgan_opt(**d)


def twostage_opt(**kw):
    """Fields TwoStageCycleModel.initialize reads (twostage_cycle_model.py:18-177); README.md:18 recipe, reduced widths via kw."""
    d = dict(isTrain=True, gpu_ids=[], checkpoints_dir="/tmp/sgk_ckpt", name="oracle", pretrained_model_dir="",
             which_channel="rg_b", batchSize=1, input_nc=2, output_nc=1, fineSize=512, norm="instance",
             noise_nc1=8, noiseSize1=4, noise_nc2=8, noiseSize2=8, ngf1=32, ngf2=64, nff2=32, ndf1=32, ndf2=64,
             which_model_netG1="fcgan", which_model_netG2="crn", which_model_netF2="unet_128",
             which_model_netD1="n_layers", which_model_netD2="n_layers", which_model_netD="n_layers",
             n_layers_G1=5, n_layers_G2=5, n_layers_F2=5, no_dropout1=True, no_dropout2=True, use_residual2=False,
             add_gaussian_noise=False, gaussian_sigma=0.1, upsample_mode1="convt", upsample_mode2="bilinear",
             n_layers_CRN_block1=1, n_layers_CRN_block2=2, no_share_label_block_weights1=False,
             no_share_label_block_weights2=False, transform_1to2="bilinear_2", scale_factor1=[1, 2], lambda_D1=[0.5, 0.4],
             n_layers_D1=[3, 3], scale_factor2=[1, 1, 2, 2], lambda_D2=[0.3, 0.3, 0.2, 0.2], n_layers_D2=[3, 4, 3, 4],
             no_lsgan1=True, no_lsgan2=True, no_cgan=False, use_multi_class_GAN=False, use_fixed_noise1=False,
             noise_pool_size=100, sequential_train=False, which_model_to_load=[""], which_epoch_sequential="seq",
             continue_train=False, which_epoch="latest", pool_size=50, lr=2e-4, lr1=2e-4, lr2=2e-4, beta1=0.5,
             which_direction="AtoB", dataset_mode="single", n_update_D1=1, n_update_D2=1, n_update_G=1,
             no_logD_trick=False, detach_G1_from_G2_x=False, detach_G1_from_G2_y=False, GAN_losses_D2=["real_fake"],
             GAN_losses_G2=["real_fake"], weights=None, lambda_A=10.0, lambda_B=10.0, lambda_A_cycle=5.0,
             lambda_fake_cycle=1.0, niter_decay=100)
    d.update(kw)
    d["scale_factor1"] = [sf(int(s)) for s in d["scale_factor1"]]
    d["scale_factor2"] = [sf(int(s)) for s in d["scale_factor2"]]
    return argparse.Namespace(**d)
