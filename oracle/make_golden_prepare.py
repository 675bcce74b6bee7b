"""Pins the oracle's data-preparation restatement (load_constraints) against the UNMODIFIED reference function and writes
the checksums to tests/golden/prepare.json.

Run in the container that has /root/reference:   python oracle/make_golden_prepare.py
  reference: processdata/PrepareData_linear.py::loadBothConstraints (:48-103), imported with the four missing third-party
  modules stubbed (pyrootutils, pytorch_lightning, cooler, matplotlib -- none is used by the function)."""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile
import types
from pathlib import Path

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF))

import numpy as np  # noqa: E402

from oracle import hicdiff_oracle as O  # noqa: E402

CASES = [dict(n_bins=150, res=40000, seed=3), dict(n_bins=700, res=40000, seed=4), dict(n_bins=333, res=10000, seed=5)]


def main():
    for m in ("pyrootutils", "pytorch_lightning", "cooler", "matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["pyrootutils"].setup_root = lambda **k: str(REF)
    sys.modules["pytorch_lightning"].LightningDataModule = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    from processdata.PrepareData_linear import loadBothConstraints

    out = {"numpy": np.__version__, "cases": []}
    for c in CASES:
        a = O.synthetic_contacts(c["n_bins"], c["res"], c["seed"])
        b = O.synthetic_contacts(c["n_bins"] + 2, c["res"], c["seed"] + 100)      # the raw-count dump: only widens the bin range
        b[:, 2] = np.round(b[:, 2])
        with tempfile.TemporaryDirectory() as td:
            fa, fb = os.path.join(td, "a.txt"), os.path.join(td, "b.txt")
            np.savetxt(fa, a)
            np.savetxt(fb, b)
            ref = loadBothConstraints(fa, fb, c["res"])
            ora = O.load_constraints(np.loadtxt(fa), np.loadtxt(fb), c["res"])
        assert ref.dtype == ora.dtype == np.float32 and ref.shape == ora.shape, (ref.dtype, ora.dtype, ref.shape, ora.shape)
        assert np.array_equal(ref, ora), f"oracle load_constraints differs from the reference for {c}"
        out["cases"].append({**c, "shape": list(ref.shape), "sha256": hashlib.sha256(np.ascontiguousarray(ref).tobytes()).hexdigest()})
        print(f"{c}: oracle == reference bit-for-bit, matrix {ref.shape}")
    (ROOT / "tests" / "golden" / "prepare.json").write_text(json.dumps(out, indent=1))
    print("wrote tests/golden/prepare.json")


if __name__ == "__main__":
    main()
