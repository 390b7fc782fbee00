from .synthetic import make_batch
