"""Path T host side: per-hospital ResNet-18 step + FedAvg over the primia_b200 C ABI."""
from .resnet18 import ResNet18Engine  # noqa: F401
from .federated import HospitalWorker, aggregation, federated_round  # noqa: F401
