from .model import CNTM
from .parameters import CNTMParameters

__all__ = ["CNTM", "CNTMParameters"]
