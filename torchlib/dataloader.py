"""torchlib/dataloader.py names used by inference.py (the image-file loaders are out of scope: SURVEY.md section 2)."""
from primia_b200.sy import RemoteTensorDataset  # noqa: F401
