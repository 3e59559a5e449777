"""torchlib/models.py of the reference -> primia_b200.models (same constructor, same state_dict keys)."""
from primia_b200.models import BasicBlock, ResNet, resnet18  # noqa: F401
