from .decoding import GreedyDecoder  # noqa: F401
