"""Top-level name `permuto_sdf` for code written against the reference (`from permuto_sdf import Sphere, ...`,
permuto_sdf_py/train_permuto_sdf.py:21-27). Put <repo>/compat and <repo> on sys.path."""
from permuto_sdf_b200.permuto_sdf import *  # noqa: F401,F403
from permuto_sdf_b200.permuto_sdf import __all__  # noqa: F401
