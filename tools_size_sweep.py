"""Diagnostic: per-call device time of gorp_extract_text_device over input sizes (reps of the 1 M-line block).
usage: python tools_size_sweep.py REPS[:ALLOC_REPS] ..."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
from gorp_b200 import _ffi, corpus
from gorp_b200.api import Blob, _check
lib = _ffi.lib
dev = torch.device("cuda", 0)
block = corpus.readme_corpus(1000000, seed=0x5EED0002)
d_block = torch.from_numpy(block.view(np.int16)).to(dev)
blob = Blob.from_definition(corpus.README_DEF)
eng = C.c_void_p(); devs = (C.c_int * 1)(0)
_check(lib.gorp_engine_create(blob._ptr, blob.length, devs, 1, C.byref(eng)))
stream = torch.cuda.current_stream().cuda_stream
dres = _ffi.DeviceResult()
for arg in sys.argv[1:]:
    reps, alloc = (int(x) for x in arg.split(":")) if ":" in arg else (int(arg), int(arg))
    d_all = d_block.repeat(alloc)
    d_text = d_all[: reps * d_block.numel()]
    n_units = d_text.numel()
    ts = []
    flush = torch.empty(64 << 20, dtype=torch.int32, device=dev) if os.environ.get("SWEEP_FLUSH") else None
    for i in range(int(os.environ.get('SWEEP_CALLS', '6'))):
        if flush is not None:
            flush.fill_(i)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _check(lib.gorp_extract_text_device(eng, 0, d_text.data_ptr(), n_units, stream, int(os.environ.get("SWEEP_FLAGS", "0")), C.byref(dres)))
        e1.record()
        if not os.environ.get("SWEEP_NOSYNC"):
            torch.cuda.synchronize()
        ts.append((e0, e1))
    torch.cuda.synchronize()
    ts = [round(a.elapsed_time(b), 2) for a, b in ts]
    print("%3d M lines (alloc %3d)" % (reps, alloc), "ms per call:", ts, "x%.2f of 0.125 ms/Mline" % (min(ts[1:]) / (reps * 0.125)), flush=True)
    del d_text, d_all
