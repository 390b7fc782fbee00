from .network import PNTransformer
from .network_refine import PNTransformerRefine
from .transformer import TransformerEncoder
