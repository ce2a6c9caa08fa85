"""Host-side logic that needs no GPU (`-m "not gpu"`).

* ImagePool: the decision plan our pool draws (image_pool.ImagePool._draw) applied by a numpy restatement of the
  sgk_image_pool_query kernel's contract must return exactly the batches the UNMODIFIED reference pool returns
  (util/image_pool.py:13-33) under the same Python `random` state -- this is what makes the captured-graph pool a
  drop-in.  The restatement is also checked against a plain-Python replay of the reference algorithm so the test still
  says something where the reference tree is absent.
* data.draw_params: the crop / flip / rotation draws consume Python `random` in the order of the reference's transform
  list (data/base_dataset.py:17-55).
* ConvCfg geometry: output sizes of Conv2d / ConvTranspose2d (with output_padding) as torch computes them.
"""
import importlib.util
import os
import random

import numpy as np
import pytest
import torch


def apply_plan(pool, plan, batch):
    """Contract of sgk_image_pool_query (include/sgk.h): walk the batch in order; -1 passes the image through,
    2*slot stores it in the pool and returns it, 2*slot+1 returns the pool's image and stores the new one."""
    out = np.empty_like(batch)
    for i, code in enumerate(plan):
        if code < 0:
            out[i] = batch[i]
        elif code % 2 == 0:
            pool[code // 2] = batch[i]
            out[i] = batch[i]
        else:
            out[i] = pool[code // 2]
            pool[code // 2] = batch[i]
    return out


def reference_pool_replay(state, pool_size, reject, batch):
    """Plain-Python replay of util/image_pool.py:13-33 on numpy arrays (state = [num_imgs, list of images])."""
    out = []
    for img in batch:
        if state[0] < pool_size:
            state[0] += 1
            state[1].append(img.copy())
            out.append(img)
        else:
            if random.uniform(0, 1) > reject:
                j = random.randint(0, pool_size - 1)
                out.append(state[1][j].copy())
                state[1][j] = img.copy()
            else:
                out.append(img)
    return np.stack(out)


def _load_reference_pool():
    from oracle import ref_loader
    path = os.path.join(ref_loader.REF_ROOT, "util", "image_pool.py")
    if not os.path.isfile(path):
        return None
    spec = importlib.util.spec_from_file_location("_ref_image_pool", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ImagePool


@pytest.mark.parametrize("pool_size,reject,B", [(5, 0.5, 3), (50, 0.5, 8), (4, 0.1, 1), (7, 0.9, 16), (2, 0.5, 2)])
def test_image_pool_decisions_match_reference(pool_size, reject, B):
    from supervised_gan_b200.image_pool import ImagePool
    RefPool = _load_reference_pool()
    rng = np.random.RandomState(pool_size * 131 + B)
    batches = [rng.randn(B, 2, 3, 3).astype(np.float32) for _ in range(40)]

    random.seed(1234)
    ours = ImagePool(pool_size, reject)
    store = np.zeros((pool_size, 2, 3, 3), np.float32)
    got = [apply_plan(store, ours._draw(B), b) for b in batches]
    end_state = random.getstate()

    random.seed(1234)
    st = [0, []]
    want = [reference_pool_replay(st, pool_size, reject, b) for b in batches]
    assert random.getstate() == end_state, "our pool consumed a different number of random draws"
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)

    if RefPool is None:
        return
    random.seed(1234)
    ref = RefPool(pool_size, reject)
    for g, b in zip(got, batches):
        r = ref.query(torch.from_numpy(b))
        np.testing.assert_array_equal(g, r.numpy())
    assert random.getstate() == end_state


def test_image_pool_size_zero_is_identity_and_draws_nothing():
    from supervised_gan_b200.image_pool import ImagePool
    random.seed(7)
    before = random.getstate()
    p = ImagePool(0)
    assert p._draw(4) == [-1] * 4
    assert random.getstate() == before
    x = torch.zeros(2, 1, 2, 2)
    assert p.query(x) is x                  # the reference returns its argument (image_pool.py:14-15)


def test_image_pool_rejects_cpu_tensors():
    from supervised_gan_b200.image_pool import ImagePool
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ImagePool(3).query(torch.zeros(2, 1, 2, 2))


def test_draw_params_consumes_random_like_the_reference_transform_list():
    """RandomCrop draws the row offset, then the column offset (torchvision's get_params: i = randint(0, h - th),
    j = randint(0, w - tw)), RandomHorizontalFlip one random() < 0.5, the rotation one randint(0, 3) -- in the list order
    of data/base_dataset.py:17-55; nothing is drawn for a switched-off transform."""
    from supervised_gan_b200 import data as D
    import argparse
    mk = lambda **kw: argparse.Namespace(**{**dict(resize_or_crop="resize_and_crop", loadSize=20, fineSize=16, no_flip=False,
                                                   no_rotate=False, isTrain=True), **kw})
    random.seed(3)
    a = D.draw_params(mk(), 20, 20)
    random.seed(3)
    b = D.draw_params(mk(), 20, 20)
    assert a == b
    random.seed(3)
    y0, x0 = random.randint(0, 4), random.randint(0, 4)
    flip = 1 if random.random() < 0.5 else 0
    rot = random.randint(0, 3)
    assert a == (y0, x0, flip, rot)
    random.seed(3)
    s0 = random.getstate()
    c = D.draw_params(mk(no_flip=True, no_rotate=True, resize_or_crop="none"), 16, 16)
    assert random.getstate() == s0, "no random decision may be drawn when crop / flip / rotation are all off"
    assert c == (0, 0, 0, 0)
    # torchvision's RandomCrop.get_params draws nothing when the image already has the crop size
    c = D.draw_params(mk(no_flip=True, no_rotate=True, loadSize=16), 16, 16)
    assert random.getstate() == s0 and c == (0, 0, 0, 0)
    with pytest.raises(ValueError):
        D.draw_params(mk(), 12, 20)


@pytest.mark.parametrize("k,s,p,H,tr,op", [(4, 2, 2, 512, 0, 0), (4, 1, 2, 65, 0, 0), (4, 2, 1, 8, 1, 0), (3, 2, 1, 64, 1, 1),
                                           (7, 1, 0, 70, 0, 0), (3, 1, 1, 33, 0, 0), (4, 1, 0, 1, 1, 0)])
def test_conv_cfg_geometry_matches_torch(k, s, p, H, tr, op):
    import supervised_gan_b200 as S
    cfg = S.ops.ConvCfg(tr, k, s, p, op)
    x = torch.zeros(1, 4, H, H + 3)
    if tr:
        w = torch.zeros(4, 8, k, k)
        y = torch.nn.functional.conv_transpose2d(x, w, stride=s, padding=p, output_padding=op)
    else:
        w = torch.zeros(8, 4, k, k)
        y = torch.nn.functional.conv2d(x, w, stride=s, padding=p)
    d = cfg.desc((1, H, H + 3, 4), w)
    assert (d.N, d.Cin, d.Cout, d.Hout, d.Wout) == (1, 4, 8, y.shape[2], y.shape[3])
    with pytest.raises(RuntimeError, match="channels"):
        cfg.desc((1, H, H + 3, 5), w)


def test_conv_cfg_rejects_output_padding_outside_transposed():
    import supervised_gan_b200 as S
    with pytest.raises(NotImplementedError):
        S.ops.ConvCfg(0, 3, 2, 1, 1)
    with pytest.raises(NotImplementedError):
        S.ops.ConvCfg(1, 3, 2, 1, 2)
