"""tfhe_fbs_map_b200 -- B200-native encrypted executor for tfhe_fbs_map's mapped circuits.

Public surface mirrors the reference (ssmiler/tfhe_fbs_map, fbs_mapper/): ``LutExecEnv`` (alias ``FbsExecEnv``),
``BitExecEnv``; evaluation runs on the GPU through the C ABI in include/fbs_b200.h.
"""
from .lut_env import LutExecEnv, FbsExecEnv, B200FbsExecEnv
from .bit_env import BitExecEnv
from .levelize import levelize, Program, table_mode, min_fbs_size
from . import params
from .mapper import MapToFBSBasic, MapToFBSHeur

__all__ = ["LutExecEnv", "FbsExecEnv", "B200FbsExecEnv", "BitExecEnv", "levelize", "Program", "table_mode",
           "min_fbs_size", "params", "MapToFBSBasic", "MapToFBSHeur"]
