from .RBF_mb import RBF_mb  # noqa: F401
from .Dot_mb import Dot_mb  # noqa: F401
