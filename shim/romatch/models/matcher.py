"""`romatch.models.matcher` of the shim: the matcher class callers type-check against (reference: `matcher.py:550`)."""
from roma_b200.matcher import RegressionMatcher  # noqa: F401
