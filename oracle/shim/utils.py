"""Stub of the reference's top-level `utils` module: the samplers import only `show_image`."""


def show_image(*a, **k):
    pass
