"""Generates tests/golden/phoc_golden.npz with the REFERENCE's own PHOC code: oracle/build_ref.py compiles
/root/reference/pythia/utils/phoc/src/cphoc.c in place into oracle/_ref/cphoc.so and binds it to the
reference's python wrapper (pythia/utils/phoc/build_phoc.py).  Runs only in the dev container; the fixture is
committed and travels.

    python tests/golden/make_phoc_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402
from vitxt_gqa_b200 import synth  # noqa: E402


def main():
    ref_build_phoc = build_ref.load_reference_build_phoc()
    assert ref_build_phoc is not None, "needs /root/reference"
    tokens = synth.make_ocr_tokens(1500, seed=2024)
    rows = np.stack([ref_build_phoc(t) for t in tokens])
    assert rows.shape == (len(tokens), 604) and rows.dtype == np.float32
    assert set(np.unique(rows).tolist()) <= {0.0, 1.0}
    out = os.path.join(ROOT, "tests", "golden", "phoc_golden.npz")
    np.savez_compressed(out, tokens_utf8=np.frombuffer("\x00".join(tokens).encode("utf-8"), np.uint8),
                        bits=np.packbits(rows.astype(np.uint8), axis=1), seed=2024)
    print("wrote", out, rows.shape, "ones per row: %.2f" % rows.sum(1).mean())


if __name__ == "__main__":
    main()
