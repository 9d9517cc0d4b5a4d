from .state import *  # noqa: F401,F403
from .goal import *  # noqa: F401,F403
from .helpers import *  # noqa: F401,F403
from .nodegoal import *  # noqa: F401,F403
from .edgegoal import *  # noqa: F401,F403
from .networkgoal import *  # noqa: F401,F403

__all__ = [name for name in dir() if not name.startswith("_")]
