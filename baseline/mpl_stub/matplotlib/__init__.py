"""Two-line stand-in for matplotlib so that the UNMODIFIED reference package under baseline/_ref imports on a
box without matplotlib (raytracingGRFF/build_rays.py:15-17 imports it at module top; nothing on the timed
path plots).  Used only by bench.py's reference timing leg and tests/golden/make_golden.py."""


def use(*args, **kwargs):
    pass
