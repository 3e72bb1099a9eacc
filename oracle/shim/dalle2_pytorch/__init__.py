"""Shim so that /root/reference/dalle2_video/dalle2_video.py:13-14 can be imported unmodified.
Test infrastructure only; see oracle/primitives.py."""
from oracle.primitives import __version__  # noqa: F401
