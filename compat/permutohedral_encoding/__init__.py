"""Top-level name `permutohedral_encoding` (`import permutohedral_encoding as permuto_enc`,
permuto_sdf_py/models/models.py:20). Put <repo>/compat and <repo> on sys.path."""
from permuto_sdf_b200.permutohedral_encoding import *  # noqa: F401,F403
from permuto_sdf_b200.permutohedral_encoding import __all__  # noqa: F401
