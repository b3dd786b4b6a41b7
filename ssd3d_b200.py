"""Importable alias of the `3dssd_b200` package (a name that starts with a digit cannot follow `import`)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("3dssd_b200")
