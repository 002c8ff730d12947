"""Generate tests/golden/dataio/: files written / read by THE REFERENCE'S OWN Python code, for pinning mc-cnn_b200/dataio.py.

Run in the build container (needs /root/reference; the fixtures are committed, the GPU box never reads the reference):

    python oracle/make_dataio_golden.py

* `tofile` is exec'd from /root/reference/preprocess_mb.py:99-106 (the module itself cannot be imported: its top level
  parses sys.argv and scans dataset directories) and writes the `<name>`, `<name>.dim`, `<name>.type` triples main.lua reads;
* `load_pfm` / `save_pfm` come from preprocess_mb.py:13-74 (Python-2 text-mode I/O, run through 2to3-style fixes noted below);
* the `.bin` layout of `-a predict` is pinned the way samples/load_bin.py reads it: np.memmap(float32, (1, D, H, W)).
"""
import ast
import json
import os
import sys

import numpy as np

REF = "/root/reference/preprocess_mb.py"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "dataio")


def reference_functions(names):
    src = open(REF).read()
    tree = ast.parse(src)
    ns = {"np": np, "os": os, "re": __import__("re"), "sys": sys}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = ast.get_source_segment(src, node)
            exec(compile(code, REF, "exec"), ns)
    return [ns[n] for n in names]


def main():
    os.makedirs(OUT, exist_ok=True)
    (tofile,) = reference_functions(["tofile"])
    rng = np.random.default_rng(0)
    arrays = {
        "x_f32": rng.standard_normal((2, 1, 5, 7)).astype(np.float32),     # x0 / x1 batches (preprocess_mb.py:171-176)
        "te_i32": np.arange(1, 7, dtype=np.int32),                          # te.bin index lists
        "meta_i32": np.array([[5, 7, 60], [6, 8, 70]], dtype=np.int32),     # meta.bin (height, width, ndisp)
        "none": None,                                                       # missing ground truth: '.dim' = '0'
    }
    for name, a in arrays.items():
        tofile(os.path.join(OUT, name + ".bin"), a)
    np.savez(os.path.join(OUT, "expected.npz"), **{k: v for k, v in arrays.items() if v is not None})
    # a PFM the way the reference's own writer lays it out would need Python 2 (file.write of str + tofile on a text-mode
    # handle); the reader is pure parsing, so pin OUR reader against the reference's on a file written with plain numpy
    img = rng.standard_normal((4, 6)).astype("<f4")
    with open(os.path.join(OUT, "disp.pfm"), "wb") as f:
        f.write(b"Pf\n6 4\n-0.003922\n")
        np.flipud(img).tofile(f)                                            # PFM rows run bottom-up
    (load_pfm,) = reference_functions(["load_pfm"])
    # the reference opens the file in text mode (Python 2); give it a binary handle with the same interface
    import builtins

    real_open = builtins.open

    class _F:
        def __init__(self, fname):
            self.f = real_open(fname, "rb")

        def readline(self):
            return self.f.readline().decode("ascii")

        def fileno(self):
            return self.f.fileno()

        def __getattr__(self, k):
            return getattr(self.f, k)

    load_pfm.__globals__["open"] = lambda fname, *a: _F(fname)
    load_pfm.__globals__["np"] = type("npshim", (), {"fromfile": staticmethod(lambda fh, dt: np.fromfile(fh.f, dt)),
                                                     "flipud": staticmethod(np.flipud), "reshape": staticmethod(np.reshape)})
    got, scale = load_pfm(os.path.join(OUT, "disp.pfm"), False)
    np.save(os.path.join(OUT, "disp_pfm_as_read_by_reference.npy"), np.ascontiguousarray(got, dtype=np.float32))
    json.dump({"pfm_scale": scale, "generated_by": "oracle/make_dataio_golden.py from /root/reference/preprocess_mb.py (tofile :99-106, load_pfm :13-57)"},
              open(os.path.join(OUT, "README.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
