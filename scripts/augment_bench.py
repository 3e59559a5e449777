"""GPU augmentation front end: one batch of 64 raw 1024x1024 uint8 images -> fp32 [64,3,224,224]; wall time of GpuAugment.apply
(host packing + 64 MB H2D + kernel) and of the kernel alone (CUDA events around a second launch on resident data)."""
import sys, time, types
import numpy as np, torch
sys.path.insert(0, '.')
from primia_b200.train.augment import GpuAugment
from primia_b200 import _lib

args = types.SimpleNamespace(rotation=30, translate=0.0, scale=0.15, shear=10, inference_resolution=512, train_resolution=224,
                             albu_prob=0.75, individual_albu_probs=0.2, noise_std=0.05, noise_prob=0.5, pretrained=True)
rng = np.random.default_rng(0)
for hw, clahe in ((1024, False), (1000, False), (1000, True)):
    args.clahe = clahe
    aug = GpuAugment(args, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225], "cuda:0", seed=0)
    images = [rng.integers(0, 256, (hw, hw), dtype=np.uint8) for _ in range(64)]
    aug(images); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        x = aug(images)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 5
    # kernel alone: replay the last launch arguments through a CUDA graph-free loop with events
    real_call = _lib.call
    saved = {"calls": []}
    def spy(name, *a):
        if name in ("pm_augment_batch_u8_f32", "pm_clahe_luts_u8", "pm_augment_clahe_finish_f32"):
            saved["calls"].append((name, a))
        return real_call(name, *a)
    import primia_b200.train.augment as G
    G.call = spy
    keep = aug(images); torch.cuda.synchronize()
    G.call = real_call
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        for name, a in saved["calls"]:
            real_call(name, *a)
    e1.record(); torch.cuda.synchronize()
    k = e0.elapsed_time(e1) / 20
    out_bytes = 64 * 3 * 224 * 224 * 4
    print(f"source {hw}x{hw} clahe={clahe}: apply() {wall * 1e3:.2f} ms/batch ({64 / wall:.0f} images/s incl. host packing + H2D); kernels {k * 1e3:.1f} us "
          f"({out_bytes / k / 1e6:.0f} GB/s of output)")
