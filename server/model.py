"""Drop-in replacement for the reference's ``server/model.py``: ``from model import run`` (server/server.py:35)
keeps working unchanged when this file replaces the reference's; the work happens in libtruely_b200.so."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import truely_b200  # noqa: E402,F401
from truely_b200.model import run  # noqa: E402,F401
