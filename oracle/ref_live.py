"""TEST INFRASTRUCTURE -- the live reference, as the oracle's pin (see baseline/ref_loader.py for the loader).

`oracle/make_golden.py` and the CPU tests import the unmodified reference through this name; on the GPU box the same
loader resolves to the shipped copy under baseline/_ref.
"""
from baseline.ref_loader import (REFERENCE_DIR, available, build_model, kind, load)  # noqa: F401
