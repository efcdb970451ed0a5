"""Drop-in import shim: `import romatch` resolves to the B200-native implementation (roma_b200).

Put this directory on PYTHONPATH *instead of* the reference checkout:

    PYTHONPATH=/path/to/repo/shim:/path/to/repo python demo/demo_match.py

Mirrors the import surface of the reference package root (`romatch/__init__.py:2-8`): the three factories and the module
globals some callers read (`romatch.RANK` gates tqdm bars, `DEBUG_MODE`, `GLOBAL_STEP`, `STEP_SIZE`, `LOCAL_RANK`).
Only the dense match()/sample() inference path exists behind it; training, datasets and benchmarks do not.
"""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.append(_ROOT)

from roma_b200 import roma_indoor, roma_outdoor, tiny_roma_v1_outdoor  # noqa: E402,F401

DEBUG_MODE = False
RANK = int(_os.environ.get("RANK", default=0))
GLOBAL_STEP = 0
STEP_SIZE = 1
LOCAL_RANK = -1
