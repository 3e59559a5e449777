"""GPU front end of the training loader: pm_augment_batch_u8_f32 against the augmentation oracle, which
tests/test_oracle_augment.py pins bit for bit to PIL / OpenCV / torchvision (torchlib/dataloader.py:138-217)."""
import types

import numpy as np
import pytest
import torch

from oracle import augment_oracle as A

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def _args(**kw):
    d = dict(rotation=30, translate=0.05, scale=0.15, shear=10, inference_resolution=512, train_resolution=224, albu_prob=0.75,
             individual_albu_probs=0.2, noise_std=0.05, noise_prob=0.5, pretrained=True)
    d.update(kw)
    return types.SimpleNamespace(**d)


def _oracle(im, p, R, T, cout, clahe=False):
    H, W = im.shape[:2]
    m = A.inverse_affine_matrix([W * 0.5, H * 0.5], p["angle"], p["translate"], p["scale"], p["shear"])
    if clahe and cout == 3:
        from primia_b200.train import _lab_tables as tab

        u8, f = A.restated_pipeline(im, m, R, T, p["cy"], p["cx"], p["flip"], MEAN, STD, clahe=True,
                                    rgb_tables=(list(tab.GREY_TO_L), list(tab.L_TO_RGB)))
        return np.ascontiguousarray(u8.transpose(2, 0, 1)), f
    u8, f = A.restated_pipeline(im, m, R, T, p["cy"], p["cx"], p["flip"], MEAN, STD, clahe=clahe)
    if u8.ndim == 2:
        u8 = u8[:, :, None]
    u8 = np.ascontiguousarray(u8.transpose(2, 0, 1))
    if u8.shape[0] < cout:   # a one-channel source feeding the 3-channel model: the RGB loader replicates it
        u8 = np.repeat(u8, cout, axis=0)
        f = np.stack([A.to_float_normalize(u8[c], [MEAN[c]], [STD[c]])[0] for c in range(cout)])
    return u8, f


@pytest.mark.parametrize("pretrained", [True, False])
def test_batch_equals_oracle_bit_for_bit(pretrained):
    from primia_b200.train.augment import GpuAugment

    rng = np.random.default_rng(5)
    R, T = 512, 224
    aug = GpuAugment(_args(pretrained=pretrained, noise_prob=0.0), MEAN, STD, DEV, seed=1)
    shapes = [(1024, 1024), (300, 420), (777, 512), (512, 512), (300, 420), (1100, 901), (64, 80), (1024, 1024)]
    images = []
    for i, (H, W) in enumerate(shapes):
        rgb = pretrained and i % 3 == 0
        images.append(rng.integers(0, 256, (H, W, 3) if rgb else (H, W), dtype=np.uint8))
    params = [aug.sample_params(*im.shape[:2]) for im in images]
    params[0]["flip"], params[1]["flip"] = True, False
    out, u8 = aug.apply(images, params, return_u8=True)
    torch.cuda.synchronize()
    assert out.shape == (len(images), aug.cout, T, T) and out.dtype == torch.float32
    for i, (im, p) in enumerate(zip(images, params)):
        ref_u8, ref_f = _oracle(im, p, R, T, aug.cout)
        assert np.array_equal(u8[i].cpu().numpy(), ref_u8), f"image {i} {im.shape}: uint8 stage differs"
        assert np.array_equal(out[i].cpu().numpy(), ref_f), f"image {i}: float stage differs"


@pytest.mark.parametrize("pretrained", [True, False])
def test_clahe_batch_equals_oracle_bit_for_bit(pretrained):
    """clahe = yes (the shipped pneumonia configs): cv::CLAHE between the crop and the flip, on the L plane of the LAB round trip
    for the 3-channel model -- the oracle's restatement is pinned to OpenCV in tests/test_oracle_augment.py"""
    from primia_b200._lib import PrimiaError
    from primia_b200.train.augment import GpuAugment

    rng = np.random.default_rng(15)
    R, T = 512, 224
    aug = GpuAugment(_args(pretrained=pretrained, noise_prob=0.0, clahe=True), MEAN, STD, DEV, seed=5)
    images = [np.clip(rng.normal(110, 45, (H, W)), 0, 255).astype(np.uint8) for H, W in [(1024, 1024), (640, 480), (333, 901), (512, 512)]]
    images.append(np.full((300, 300), 90, dtype=np.uint8))        # flat image: every tile's histogram is one clipped spike
    params = [aug.sample_params(*im.shape[:2]) for im in images]
    params[0]["flip"], params[1]["flip"] = True, False
    out, u8 = aug.apply(images, params, return_u8=True)
    torch.cuda.synchronize()
    for i, (im, p) in enumerate(zip(images, params)):
        ref_u8, ref_f = _oracle(im, p, R, T, aug.cout, clahe=True)
        assert np.array_equal(u8[i].cpu().numpy(), ref_u8), f"image {i} {im.shape}: uint8 stage differs"
        assert np.array_equal(out[i].cpu().numpy(), ref_f), f"image {i}: float stage differs"
    if pretrained:
        with pytest.raises(PrimiaError):
            aug.apply([rng.integers(0, 256, (300, 300, 3), dtype=np.uint8)], [params[0]])


def test_identity_parameters_reproduce_a_plain_resize_crop():
    """no rotation / shear / scale change, no flip: the kernel is cv2.resize + crop + normalize of the untouched image"""
    import cv2

    from primia_b200.train.augment import GpuAugment

    rng = np.random.default_rng(6)
    aug = GpuAugment(_args(rotation=0, translate=0.0, scale=0.0, shear=0, albu_prob=0.0, inference_resolution=256, train_resolution=256,
                           pretrained=False), MEAN, STD, DEV, seed=2)
    im = rng.integers(0, 256, (333, 450), dtype=np.uint8)
    out, u8 = aug.apply([im], [aug.sample_params(333, 450)], return_u8=True)
    assert np.array_equal(u8[0, 0].cpu().numpy(), cv2.resize(im, (256, 256), interpolation=cv2.INTER_LINEAR))


def test_gauss_noise_statistics_and_seeding():
    from primia_b200.train.augment import GpuAugment

    aug = GpuAugment(_args(rotation=0, translate=0.0, scale=0.0, shear=0, inference_resolution=224, train_resolution=224, pretrained=False),
                     MEAN, STD, DEV, seed=3)
    im = np.full((224, 224), 128, dtype=np.uint8)
    base = dict(angle=0.0, translate=(0, 0), scale=1.0, shear=(0.0, 0.0), cy=0, cx=0, flip=False)
    _, clean = aug.apply([im], [dict(base, noise_sigma=0.0, noise_seed=0)], return_u8=True)
    _, n1 = aug.apply([im], [dict(base, noise_sigma=6.0, noise_seed=11)], return_u8=True)
    _, n1b = aug.apply([im], [dict(base, noise_sigma=6.0, noise_seed=11)], return_u8=True)
    _, n2 = aug.apply([im], [dict(base, noise_sigma=6.0, noise_seed=12)], return_u8=True)
    assert torch.equal(n1, n1b) and not torch.equal(n1, n2)
    d = n1.double() - clean.double()
    # image + N(0, 6^2), clipped, truncated toward zero by the uint8 cast: mean shifts by about -0.5, spread stays ~6
    assert abs(d.mean().item() + 0.5) < 0.15 and abs(d.std().item() - 6.0) < 0.3
    k = ((d - d.mean()) ** 4).mean() / d.var() ** 2
    assert abs(k.item() - 3.0) < 0.3     # Gaussian kurtosis
    # the reference's own setting (noise_std 0.05 on a 0..255 image): at most one grey level, downwards (truncation)
    _, tiny = aug.apply([im], [dict(base, noise_sigma=0.05, noise_seed=5)], return_u8=True)
    dt = tiny.int() - clean.int()
    assert set(dt.unique().tolist()) <= {-1, 0} and 0.3 < (dt == -1).float().mean().item() < 0.7


def test_augmented_batch_feeds_the_training_engine():
    from primia_b200.train import ResNet18Engine
    from primia_b200.train.augment import GpuAugment

    rng = np.random.default_rng(8)
    aug = GpuAugment(_args(inference_resolution=128, train_resolution=96), MEAN, STD, DEV, seed=4)
    images = [rng.integers(0, 256, (int(rng.integers(100, 300)), int(rng.integers(100, 300))), dtype=np.uint8) for _ in range(4)]
    x = aug(images)
    assert x.shape == (4, 3, 96, 96) and torch.isfinite(x).all()
    eng = ResNet18Engine(4, 3, 3, 96, "max", DEV, "bf16")
    eng.init_random(1)
    loss = eng.train_step(x, torch.tensor([0, 1, 2, 1], device=DEV))
    assert torch.isfinite(loss).all()
    # reproducible parameter stream
    a1 = GpuAugment(_args(), MEAN, STD, DEV, seed=77).sample_params(600, 500)
    a2 = GpuAugment(_args(), MEAN, STD, DEV, seed=77).sample_params(600, 500)
    assert a1 == a2
