from . import linear_interpolate  # noqa: F401  (MGRU.py:4 imports the MODULE)
