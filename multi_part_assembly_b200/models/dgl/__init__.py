from .network import DGLModel
