"""Inert stand-in so the reference package imports without the `strenum` wheel (golden generation only)."""
from enum import StrEnum  # noqa: F401  (Python >= 3.11)
