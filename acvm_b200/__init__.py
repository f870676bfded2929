"""acvm_b200 -- B200-native batched ACIR witness solver (drop-in for the acvm::pwg hot path).

The product is libacvm_b200.so (hand-written sm_100a CUDA + C++ host, C ABI in include/acvm_b200.h);
this package is the thin Python mirror of that ABI.  Importing the solver without the built
library raises ImportError -- there is no CPU implementation behind it.
"""
from .solver import (ACVM, AcvmError, CompiledCircuit, Context, DeviceBatch, InstanceStatus, compile_plan_host,  # noqa: F401
                     compress_witness_map, decompress_witness_map, lib, witness_checksum)

__all__ = ["ACVM", "AcvmError", "CompiledCircuit", "Context", "DeviceBatch", "InstanceStatus", "compile_plan_host", "lib"]
