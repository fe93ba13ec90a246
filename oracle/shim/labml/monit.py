"""Stub of labml.monit (progress bars only) so the reference samplers import without labml."""


def iterate(name, it, **kw):
    return iter(range(it) if isinstance(it, int) else it)


def enum(name, it, **kw):
    return enumerate(it)
