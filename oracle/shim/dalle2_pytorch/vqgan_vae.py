"""Import target of /root/reference/dalle2_video/dalle2_video.py:14 (oracle shim)."""
from oracle.primitives import NullVQGanVAE, VQGanVAE  # noqa: F401
