"""gorp_b200 — B200-native batch extraction engine for salesforce/gorp definitions.

The product is libgorpcuda.so (gorp_b200/csrc, C ABI in include/gorp_cuda.h); this package is the thin
host-side mirror of the reference API used by tests and bench.py. Names are resolved lazily so that
`python -m gorp_b200.build` can run before the shared library exists.
"""
_API = ("Blob", "CookedExtraction", "DefinitionParseException", "DefinitionReader", "ExtractionBatch",
        "ExtractionException", "ExtractionResult", "Gorp", "GorpCudaError", "UnsupportedDefinition", "to_units")


def __getattr__(name):
    if name in _API:
        from . import api
        return getattr(api, name)
    raise AttributeError("module 'gorp_b200' has no attribute %r" % name)
