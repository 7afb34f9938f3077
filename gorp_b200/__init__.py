"""gorp_b200 — B200-native batch extraction engine for salesforce/gorp definitions.

The product is libgorpcuda.so (gorp_b200/csrc, C ABI in include/gorp_cuda.h); this package is the thin
host-side mirror of the reference API used by tests and bench.py.
"""
from .api import (Blob, CookedExtraction, DefinitionParseException, DefinitionReader, ExtractionBatch,  # noqa: F401
                  ExtractionException, ExtractionResult, Gorp, GorpCudaError, UnsupportedDefinition, to_units)
