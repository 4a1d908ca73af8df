"""Re-export of the synthetic frame generators (kept at the repo root in synth_frames.py because bench.py's
GPU arm uses them too and must not import from oracle/)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synth_frames import *  # noqa: F401,F403,E402
from synth_frames import _canvas  # noqa: F401,E402
