"""resnet_b200 -- B200-native (sm_100a) ResNet training hot path behind the C API of als244/ResNet's resnet.h.

The product is resnet_b200/libresnet_b200.so (CUDA C++, built by resnet_b200/build.py); this package is the thin
host-side mirror used by tests and bench.py.  See DESIGN.md and INTEGRATION.md."""
from . import lib  # noqa: F401
from .api import Trainer  # noqa: F401
