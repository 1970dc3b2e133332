import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Margin-aware answer-index comparisons: rows checked / mismatches outside the margin band (must be 0) /
    reference rows inside the band / samples whose free-running decode took a tolerated flip."""
    try:
        from parity_utils import PARITY_REPORT
    except Exception:
        return
    if not PARITY_REPORT:
        return
    tr = terminalreporter
    tr.write_sep("-", "answer-index parity (margin band %s on the reference's top1-top2)" % "1e-1")
    tot = dict(checked=0, mismatch=0, low_margin=0, flipped_samples=0)
    for name, f in PARITY_REPORT:
        tr.write_line("%-58s %s" % (name, "  ".join("%s=%s" % kv for kv in f.items())))
        for k in tot:
            tot[k] += int(f.get(k, 0))
    tr.write_line("TOTAL  " + "  ".join("%s=%d" % kv for kv in tot.items()))
