"""`romatch.models` of the shim: the factories (reference: `romatch/models/__init__.py`)."""
from roma_b200 import roma_indoor, roma_outdoor, tiny_roma_v1_outdoor  # noqa: F401
from roma_b200.matcher import RegressionMatcher  # noqa: F401
