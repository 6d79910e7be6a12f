from . import linear  # noqa: F401  (MGRU.py:6 imports the MODULE and calls linear.Linear)
