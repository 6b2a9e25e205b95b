from .meta_ngp import MetaNGP
from .meta_container import MetaContainer, build_expert

__all__ = ["MetaNGP", "MetaContainer", "build_expert"]
