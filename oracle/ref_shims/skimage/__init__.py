"""Stand-in for scikit-image (not installed): only what the reference imports (utils.py:14).  TEST INFRASTRUCTURE ONLY."""
from . import metrics  # noqa: F401
