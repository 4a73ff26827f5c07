import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _inert import make as _make  # noqa: E402

_m = _make("cartopy", ("crs", "feature"))
crs, feature = _m.crs, _m.feature
