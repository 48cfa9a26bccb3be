import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


import os

# drop-in entry points (cli, side_chain_builder, mcsce_sidechain) resolve the library themselves; the tests run them on
# the seeded synthetic stand-in, which production use must opt into explicitly
os.environ.setdefault("MCSCE_B200_ALLOW_SYNTHETIC", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")
