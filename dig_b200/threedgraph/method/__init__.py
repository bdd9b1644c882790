"""Same export list as reference dig/threedgraph/method/__init__.py:1-16 (ProNet is a "next" row,
SURVEY.md 8f)."""
from .dimenet_family import SphereNet, DimeNetPP
from .schnet import SchNet

__all__ = ['SchNet', 'DimeNetPP', 'SphereNet']
