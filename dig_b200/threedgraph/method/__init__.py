"""Same export list as reference dig/threedgraph/method/__init__.py:1-16."""
from .run import run
from .schnet import SchNet
from .dimenet_family import DimeNetPP, SphereNet
from .comenet import ComENet
from .pronet import ProNet

__all__ = ['run', 'SchNet', 'DimeNetPP', 'SphereNet', 'ComENet', 'ProNet']
