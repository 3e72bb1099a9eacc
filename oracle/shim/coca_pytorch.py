"""Stub: the hot path only names CoCa in an isinstance check (dalle2_video.py:1250)."""


class CoCa:  # pragma: no cover
    pass
