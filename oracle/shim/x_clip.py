"""Stub: the hot path only names CLIP in an isinstance check (dalle2_video.py:1248)."""


class CLIP:  # pragma: no cover
    pass
