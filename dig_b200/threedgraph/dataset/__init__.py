"""Readers for the 3D-graph datasets of the reference (dig/threedgraph/dataset/__init__.py:1-11), SURVEY.md 8f rank 3.
ECdataset / FOLDdataset (protein benchmarks) are not provided."""
from .datasets import MD17, QM93D

__all__ = ['QM93D', 'MD17']
