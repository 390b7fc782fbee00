from .network import GlobalModel
