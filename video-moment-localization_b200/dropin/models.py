"""``from models import SMIN`` (main.py:3) resolves here when this directory precedes the
reference on ``sys.path``."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import vml_b200  # noqa: E402,F401
from vml_b200.smin import SMIN  # noqa: E402,F401
