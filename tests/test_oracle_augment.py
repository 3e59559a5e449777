"""The augmentation oracle's integer restatement (what the CUDA kernel follows) against the REAL libraries the reference's
loader uses -- PIL's affine transform through torchvision, OpenCV's uint8 bilinear resize -- bit for bit, on random shapes and
random RandomAffine parameters (torchlib/dataloader.py:138-217)."""
import numpy as np
import pytest

from oracle import augment_oracle as A

cv2 = pytest.importorskip("cv2")
pytest.importorskip("PIL")
tv = pytest.importorskip("torchvision")


def _params(rng, W, H):
    from torchvision.transforms import RandomAffine

    # pneumonia-resnet-pretrained.ini: rotation 30, translate 0.0, scale 0.15, shear 10 (+ a non-zero translate case)
    tr = (0.1, 0.1) if rng.random() < 0.5 else None
    return RandomAffine.get_params([-30.0, 30.0], tr, (0.85, 1.15), [-10.0, 10.0], [W, H])


def test_inverse_affine_matrix_equals_torchvision():
    from torchvision.transforms.functional import _get_inverse_affine_matrix

    rng = np.random.default_rng(0)
    for _ in range(50):
        W, H = int(rng.integers(20, 2000)), int(rng.integers(20, 2000))
        angle, translate, scale, shear = _params(rng, W, H)
        ref = _get_inverse_affine_matrix([W * 0.5, H * 0.5], angle, list(translate), scale, list(shear))
        got = A.inverse_affine_matrix([W * 0.5, H * 0.5], angle, translate, scale, shear)
        assert np.allclose(ref, got, rtol=0, atol=1e-9 * max(W, H)), (ref, got)
        assert A.fix16(ref) == A.fix16(got)     # the fixed-point words PIL works with are identical


@pytest.mark.parametrize("channels", [1, 3])
def test_affine_nearest_fixed_point_equals_pil(channels):
    import torchvision.transforms.functional as TF
    from PIL import Image

    rng = np.random.default_rng(1 + channels)
    for _ in range(25):
        H, W = int(rng.integers(17, 400)), int(rng.integers(17, 400))
        shape = (H, W) if channels == 1 else (H, W, 3)
        src = rng.integers(0, 256, shape, dtype=np.uint8)
        angle, translate, scale, shear = _params(rng, W, H)
        ref = np.array(TF.affine(Image.fromarray(src), angle, list(translate), scale, list(shear)))
        m = A.inverse_affine_matrix([W * 0.5, H * 0.5], angle, translate, scale, shear)
        assert np.array_equal(A.affine_nearest_fixed(src, m), ref)


@pytest.mark.parametrize("R", [224, 256, 512, 100, 333])
def test_resize_linear_u8_equals_opencv(R):
    rng = np.random.default_rng(R)
    cases = [(int(rng.integers(20, 900)), int(rng.integers(20, 900))) for _ in range(10)] + [(2 * R, 2 * R), (R, R), (2 * R, R), (1, 7)]
    for H, W in cases:
        for shape in ((H, W), (H, W, 3)):
            src = rng.integers(0, 256, shape, dtype=np.uint8)
            ref = cv2.resize(src, (R, R), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(A.resize_linear_u8(src, R), ref), (H, W, R, shape)


def test_whole_pipeline_equals_the_libraries():
    rng = np.random.default_rng(9)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    for i in range(12):
        H, W = int(rng.integers(200, 1100)), int(rng.integers(200, 1100))
        if i == 0:
            H = W = 1024                                                  # the exact-2x box-mean path of cv::resize
        src = rng.integers(0, 256, (H, W) if i % 2 else (H, W, 3), dtype=np.uint8)
        R, T = 512, 224
        angle, translate, scale, shear = _params(rng, W, H)
        cy, cx, flip = int(rng.integers(0, R - T + 1)), int(rng.integers(0, R - T + 1)), bool(rng.integers(0, 2))
        ref_u8, ref_f = A.reference_pipeline(src, angle, translate, scale, shear, R, T, cy, cx, flip, mean, std)
        m = A.inverse_affine_matrix([W * 0.5, H * 0.5], angle, translate, scale, shear)
        got_u8, got_f = A.restated_pipeline(src, m, R, T, cy, cx, flip, mean, std)
        assert np.array_equal(ref_u8, got_u8)
        assert np.array_equal(ref_f, got_f) and got_f.dtype == np.float32 and got_f.shape[1:] == (T, T)


def test_host_side_of_the_gpu_front_end_builds_the_oracles_tables():
    """primia_b200/train/augment.py (product: never imports the oracle) derives the same fixed-point affine words and resize
    tables as the oracle, draws parameters inside the reference's ranges, and refuses transforms that are not built"""
    import types

    from primia_b200._lib import PrimiaError
    from primia_b200.train import augment as G

    rng = np.random.default_rng(0)
    for _ in range(20):
        H, W, R = int(rng.integers(20, 900)), int(rng.integers(20, 900)), int(rng.choice([224, 512, 100]))
        t = G.resize_tables(H, W, R)
        ox, oy = A.resize_tables(W, R, "x"), A.resize_tables(H, R, "y")
        assert all(np.array_equal(t[k], ox[k]) for k in range(4)) and all(np.array_equal(t[4 + k], oy[k]) for k in range(4))
        ang, tr = float(rng.uniform(-30, 30)), (int(rng.integers(-9, 9)), int(rng.integers(-9, 9)))
        sc, sh = float(rng.uniform(0.8, 1.2)), (float(rng.uniform(-10, 10)), 0.0)
        assert G.fix16(G.inverse_affine_matrix((W * 0.5, H * 0.5), ang, tr, sc, sh)) == A.fix16(
            A.inverse_affine_matrix([W * 0.5, H * 0.5], ang, tr, sc, sh))
    args = types.SimpleNamespace(rotation=30, translate=0.1, scale=0.15, shear=10, inference_resolution=512, train_resolution=224,
                                 albu_prob=0.75, individual_albu_probs=0.2, noise_std=0.05, noise_prob=0.5, pretrained=True)
    aug = G.GpuAugment(args, [0.5], [0.25], device="cpu", seed=1)
    assert aug.cout == 3 and aug.mean.tolist() == [0.5] * 3 and aug.rstd.tolist() == [4.0] * 3
    ps = [aug.sample_params(600, 800) for _ in range(2000)]
    assert all(-30 <= p["angle"] <= 30 and 0.85 <= p["scale"] <= 1.15 and -10 <= p["shear"][0] <= 10 and p["shear"][1] == 0 for p in ps)
    assert all(abs(p["translate"][0]) <= 80 and abs(p["translate"][1]) <= 60 and 0 <= p["cy"] <= 288 and 0 <= p["cx"] <= 288 for p in ps)
    flips, noisy = np.mean([p["flip"] for p in ps]), np.mean([p["noise_sigma"] > 0 for p in ps])
    assert abs(flips - 0.75 * 0.2) < 0.03 and abs(noisy - 0.75 * 0.5) < 0.04      # albu_prob gates the group, then each p
    assert all(p["noise_sigma"] <= 0.05 for p in ps)
    with pytest.raises(PrimiaError):
        G.GpuAugment(types.SimpleNamespace(**dict(vars(args), blur=True)), [0.5], [0.25], device="cpu")
    with pytest.raises(PrimiaError):   # clahe needs the crop to tile 8 x 8
        G.GpuAugment(types.SimpleNamespace(**dict(vars(args), clahe=True, train_resolution=225)), [0.5], [0.25], device="cpu")
    assert G.GpuAugment(types.SimpleNamespace(**dict(vars(args), clahe=True)), [0.5], [0.25], device="cpu").clahe


@pytest.mark.parametrize("T", [224, 256, 96, 512])
def test_clahe_restatement_equals_opencv(T):
    rng = np.random.default_rng(T)
    yy, xx = np.mgrid[0:T, 0:T]
    cases = [rng.integers(0, 256, (T, T), dtype=np.uint8), np.clip(rng.normal(120, 40, (T, T)), 0, 255).astype(np.uint8),
             ((np.sin(xx / 17.0) * 60 + np.cos(yy / 23.0) * 50 + 128) + rng.normal(0, 6, (T, T))).clip(0, 255).astype(np.uint8),
             np.full((T, T), 77, dtype=np.uint8), (xx % 256).astype(np.uint8)]
    for src in cases:
        assert np.array_equal(A.clahe_u8(src), cv2.createCLAHE(clipLimit=1.0, tileGridSize=(8, 8)).apply(src))


def test_lab_tables_are_opencvs_conversion_on_grey_pixels_and_the_rgb_recipe_follows():
    from primia_b200.train import _lab_tables as tab

    v = np.arange(256, dtype=np.uint8)
    lab = cv2.cvtColor(np.stack([v, v, v], -1)[None], cv2.COLOR_RGB2LAB)[0]
    assert (lab[:, 1:] == 128).all() and bytes(lab[:, 0].tolist()) == tab.GREY_TO_L
    rgb = cv2.cvtColor(np.stack([v, np.full(256, 128, np.uint8), np.full(256, 128, np.uint8)], -1)[None], cv2.COLOR_LAB2RGB)[0]
    assert bytes(rgb.reshape(-1).tolist()) == tab.L_TO_RGB
    rng = np.random.default_rng(3)
    for _ in range(3):
        g = np.clip(rng.normal(120, 50, (224, 224)), 0, 255).astype(np.uint8)
        ref = A.clahe_reference(np.stack([g] * 3, -1))                     # albumentations' recipe through OpenCV
        got = A.clahe_grey_rgb(g, list(tab.GREY_TO_L), list(tab.L_TO_RGB))
        assert np.array_equal(ref, got)


def test_pipeline_with_clahe_equals_the_libraries():
    from primia_b200.train import _lab_tables as tab

    rng = np.random.default_rng(10)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    for i in range(4):
        H, W = int(rng.integers(300, 1100)), int(rng.integers(300, 1100))
        g = np.clip(rng.normal(110, 45, (H, W)), 0, 255).astype(np.uint8)
        R, T = 512, 224
        angle, translate, scale, shear = _params(rng, W, H)
        cy, cx, flip = int(rng.integers(0, R - T + 1)), int(rng.integers(0, R - T + 1)), bool(i % 2)
        m = A.inverse_affine_matrix([W * 0.5, H * 0.5], angle, translate, scale, shear)
        # single-channel model
        ref_u8, ref_f = A.reference_pipeline(g, angle, translate, scale, shear, R, T, cy, cx, flip, mean, std, clahe=True)
        got_u8, got_f = A.restated_pipeline(g, m, R, T, cy, cx, flip, mean, std, clahe=True)
        assert np.array_equal(ref_u8, got_u8) and np.array_equal(ref_f, got_f)
        # 3-channel model fed by the RGB loader (grey replicated)
        ref_u8, ref_f = A.reference_pipeline(np.stack([g] * 3, -1), angle, translate, scale, shear, R, T, cy, cx, flip, mean, std, clahe=True)
        got_u8, got_f = A.restated_pipeline(g, m, R, T, cy, cx, flip, mean, std, clahe=True,
                                            rgb_tables=(list(tab.GREY_TO_L), list(tab.L_TO_RGB)))
        assert np.array_equal(ref_u8, got_u8) and np.array_equal(ref_f, got_f)
