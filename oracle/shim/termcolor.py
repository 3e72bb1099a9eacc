"""Stub for termcolor.cprint (dalle2_video.py:3)."""


def cprint(*args, **kwargs):  # pragma: no cover
    print(*args)
