from .state import *  # noqa: F401,F403
from .structure import *  # noqa: F401,F403
from .model import *  # noqa: F401,F403
from .fdm import *  # noqa: F401,F403

__all__ = [name for name in dir() if not name.startswith("_")]
