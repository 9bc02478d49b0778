"""File to list under ``custom_models:`` in a CREDIT config (credit/models/__init__.py:278-298).

CREDIT executes it before the registry lookup; it registers the B200 forecast steps under ``type: crossformer_b200``,
``wxformer_b200``, ``fuxi_b200`` and ``crossformer-ensemble_b200``.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from miles_credit_b200.model import register_with_credit  # noqa: E402

register_with_credit("crossformer_b200")
