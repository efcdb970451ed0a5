"""`romatch.models.model_zoo` of the shim (reference: `romatch/models/model_zoo/__init__.py:6-94`)."""
from roma_b200.model_zoo import roma_indoor, roma_model, roma_outdoor, tiny_roma_v1_outdoor, weight_urls  # noqa: F401
