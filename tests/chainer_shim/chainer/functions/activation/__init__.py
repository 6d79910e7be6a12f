from . import relu, sigmoid, tanh  # noqa: F401  (MGRU.py:3 imports the MODULES and calls sigmoid.sigmoid)
