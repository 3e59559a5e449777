"""Import-path compatibility with the reference: its entry points do ``from torchlib.utils import Arguments, ...``,
``from torchlib.models import resnet18`` and ``from torchlib.run_websocket_server import read_websocket_config``, and its
checkpoints pickle a ``torchlib.utils.Arguments`` object (torchlib/utils.py:1470-1493).  The modules here provide those names
on top of primia_b200, so a checkpoint written by either code base loads in the other's train.py / inference.py."""
