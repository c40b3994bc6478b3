"""Importable alias for the ``alphazero-al_b200`` package directory (a hyphen cannot appear in an import statement)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("alphazero-al_b200")
