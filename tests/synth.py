"""Test-side alias of the seeded generators (kimimaro_b200/datasets.py)."""
from kimimaro_b200.datasets import sphere, synthetic_tubes, tiled  # noqa: F401
