"""``from utils import compute_ious`` (main.py:5) resolves here when this directory precedes
the reference on ``sys.path``.  ``get_tokens`` (utils.py:6-7) is host text preparation and is
kept for import compatibility with the reference's ``dataset.py``."""
import os
import string
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import vml_b200  # noqa: E402,F401
from vml_b200.evaluate import compute_ious  # noqa: E402,F401


def get_tokens(query):
    table = str.maketrans("", "", string.punctuation)
    return str(query).lower().translate(table).strip().split()
