"""Star-import target of /root/reference/dalle2_video/dalle2_video.py:13 (oracle shim)."""
from oracle.primitives import *  # noqa: F401,F403
from oracle.primitives import __all__  # noqa: F401
