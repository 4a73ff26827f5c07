import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _inert import make as _make  # noqa: E402

_m = _make("matplotlib", ("pylab", "pyplot", "cm", "colors"))
pylab, pyplot, cm, colors = _m.pylab, _m.pyplot, _m.cm, _m.colors
