#!/usr/bin/env python3
"""Regenerate tests/golden/inputs.npz from the reference's text matrices.

Run in the build container only (needs /root/reference).  The reference ships five
whitespace-separated text matrices that every `main()` reads as a *token stream*
(e.g. /root/reference/templated/luBatchedInplace.cu:30-34): the
N x N template is `tokens[0:N*N].reshape(N, N)`.  We therefore store the token
streams, not matrices.  Two arrays per file:

  <name>_f64 : every token parsed as double  (std::ifstream >> double)
  <name>_f32 : every token parsed as float   (std::ifstream >> float == strtof)

strtof is called through libc so that the float32 array is exactly what the
reference's `file >> templateMatrix[i]` produces with `using FpType = float`
(/root/reference/templated/verify.hpp:9-10); the script asserts that it equals
float32(double) so either array can be used.
"""
import ctypes, hashlib, json, os, sys
import numpy as np

REF = "/root/reference"
FILES = {
    "mtrand32": "templated/mtrand32.txt",
    "mtrand32_new1": "templated/mtrand32_new1.txt",
    "mtrand32_new": "parallel_pivot/mtrand32_new.txt",
    "mtrand64": "templated/mtrand64.txt",
    "matrix": "templated/matrix.txt",
}

def main():
    libc = ctypes.CDLL("libc.so.6")
    libc.strtof.restype = ctypes.c_float
    libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    out, meta = {}, {}
    for name, rel in FILES.items():
        raw = open(os.path.join(REF, rel), "rb").read()
        toks = raw.split()
        f64 = np.array([float(t) for t in toks], dtype=np.float64)
        f32 = np.array([libc.strtof(t, None) for t in toks], dtype=np.float32)
        assert np.array_equal(f32, f64.astype(np.float32)), name
        out[name + "_f64"], out[name + "_f32"] = f64, f32
        meta[name] = {"source": rel, "md5": hashlib.md5(raw).hexdigest(), "tokens": len(toks)}
        print(name, len(toks), meta[name]["md5"])
    here = os.path.dirname(os.path.abspath(__file__))
    np.savez_compressed(os.path.join(here, "inputs.npz"), **out)
    json.dump(meta, open(os.path.join(here, "inputs.meta.json"), "w"), indent=1, sort_keys=True)

if __name__ == "__main__":
    sys.exit(main())
