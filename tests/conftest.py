import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# update_parameters keeps the reference's checkpoint cadence (a full dump whenever cur_dump_id % 1000 == 0, i.e. at the very first
# step of a run that uses load_new_batch) and dumps on NaN; tests that are not about dumps switch both off instead of writing
# gigabytes under the default dump root.  tests/test_gpu_c_driver.py and tests/test_gpu_dump.py set their own values.
os.environ.setdefault("RESNET_B200_DUMP_EVERY", "0")
os.environ.setdefault("RESNET_B200_DUMP_ON_NAN", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
