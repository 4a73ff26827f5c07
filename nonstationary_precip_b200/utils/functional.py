"""utils/functional.py of the reference: the four helpers the hot path uses (op, dot, mv, t; reference
utils/functional.py:14-33,60-64), backed by the npgp kernels."""
from ..functional import dot, mv, op, t  # noqa: F401
