"""torchlib/dataloader.py names used by train.py / inference.py.  The image-FILE loaders (folders, DICOM) are out of scope
(SURVEY.md section 2); the augmentation they feed -- ``create_albu_transform`` (dataloader.py:138-217) -- is built as a
batch-level GPU transform (primia_b200/train/augment.py)."""
from primia_b200.sy import RemoteTensorDataset  # noqa: F401
from primia_b200.train.augment import GpuAugment


def create_albu_transform(args, mean, std, device="cuda:0", seed=None):
    """dataloader.py:138-217.  The reference returns a per-image CPU transform (torchvision RandomAffine + albumentations);
    this returns a callable over a LIST of raw uint8 images that produces the augmented, normalised fp32 [B, C, T, T] batch on
    ``device`` in one kernel launch, with the same fixed-point arithmetic as PIL and OpenCV (tests/test_augment_gpu.py)."""
    return GpuAugment(args, mean, std, device, seed)
