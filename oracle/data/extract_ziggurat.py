"""Where oracle/data/ziggurat_normal.npz comes from: the three 256-entry tables of numpy's ziggurat normal sampler
(numpy/random/src/distributions/ziggurat_constants.h: wi_double, ki_double, fi_double) read out of the static library the
installed numpy ships (numpy/random/lib/libnpyrandom.a, object distributions.c.o, read-only data).  They are constants of a
published algorithm (Marsaglia & Tsang's ziggurat with 256 strips, r = 3.6541528853610088) in the exact doubles numpy
uses -- recomputing them from r and v in double precision does not reproduce the last digits, so they are kept as data.
Run once, in the build container:  python oracle/data/extract_ziggurat.py
"""
import os
import subprocess
import tempfile

import numpy as np

lib = os.path.join(os.path.dirname(np.__file__), "random", "lib", "libnpyrandom.a")
with tempfile.TemporaryDirectory() as tmp:
    subprocess.check_call(["ar", "x", lib], cwd=tmp)
    obj = [f for f in os.listdir(tmp) if "distributions.c" in f and "distributions_distributions" in f or f == "distributions.c.o"][0]
    path = os.path.join(tmp, obj)
    sec = [ln.split() for ln in subprocess.run(["objdump", "-h", path], capture_output=True, text=True).stdout.splitlines()
           if len(ln.split()) > 5 and ln.split()[1] == ".rodata"][0]
    base = int(sec[5], 16)
    syms = {}
    for ln in subprocess.run(["nm", "-S", path], capture_output=True, text=True).stdout.splitlines():
        p = ln.split()
        if len(p) == 4 and p[3] in ("wi_double", "ki_double", "fi_double"):
            syms[p[3]] = (int(p[0], 16), int(p[1], 16))
    raw = open(path, "rb").read()
    out = {}
    for name, dt in (("wi_double", "<f8"), ("ki_double", "<u8"), ("fi_double", "<f8")):
        off, size = syms[name]
        assert size == 2048
        out[name[:2]] = np.frombuffer(raw[base + off: base + off + size], dtype=dt).copy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ziggurat_normal.npz"), **out)
print({k: (v[:2], v[-1]) for k, v in out.items()})
