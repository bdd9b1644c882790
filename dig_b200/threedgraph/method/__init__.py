"""Same export list as reference dig/threedgraph/method/__init__.py:1-16 (ProNet is a "next" row,
SURVEY.md 8f)."""
from .run import run
from .schnet import SchNet
from .dimenet_family import DimeNetPP, SphereNet
from .comenet import ComENet

__all__ = ['run', 'SchNet', 'DimeNetPP', 'SphereNet', 'ComENet']
