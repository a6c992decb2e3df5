"""Empty pyplot stand-in (see matplotlib/__init__.py in this directory)."""
