from .model import CNVM
from .parameters import CNVMParameters, convert_rate_from_cnvm, convert_rate_to_cnvm, save_params_as_txt_file

__all__ = ["CNVM", "CNVMParameters", "convert_rate_to_cnvm", "convert_rate_from_cnvm", "save_params_as_txt_file"]
