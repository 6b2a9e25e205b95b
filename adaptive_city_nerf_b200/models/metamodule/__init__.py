from .metamodule import MetaModule, MetaSequential, MetaLinear, MetaBatchLinear, MetaLayerBlock

__all__ = ["MetaModule", "MetaSequential", "MetaLinear", "MetaBatchLinear", "MetaLayerBlock"]
