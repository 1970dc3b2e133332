"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's PHOC descriptor.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(vitxt_gqa_b200/) never does.

Follows, line by line:
  * pythia/utils/phoc/build_phoc.py:9-14   -- lower(), strip(), drop every character outside [a-z0-9]
  * pythia/utils/phoc/src/cphoc.c:12-113   -- the 604-d descriptor: unigram levels 2..5 (36 symbols x 14
    regions = 504) + the 50 most frequent English bigrams at level 2 (100)
  * pythia/datasets/processors.py:904-928  -- PhocProcessor: [max_length, 604] fp32, rows >= len(tokens) = 0

Pinning: oracle/build_ref.py compiles the reference's own cphoc.c (where it lies under /root/reference) into
oracle/_ref/cphoc.so; tests/golden/make_phoc_golden.py ran that binary through the reference's build_phoc.py
wrapper and committed tests/golden/phoc_golden.npz; tests/test_oracle_cpu.py holds this file to it bit for bit.

All arithmetic is IEEE binary32 exactly as gcc emits it for the C source (FLT_EVAL_METHOD 0 on x86-64): every
division / subtraction below is rounded to float32 once.
"""
import numpy as np

UNIGRAMS = "abcdefghijklmnopqrstuvwxyz0123456789"                      # cphoc.c:29
BIGRAMS = ["th", "he", "in", "er", "an", "re", "es", "on", "st", "nt", "en", "at", "ed", "nd", "to", "or", "ea",
           "ti", "ar", "te", "ng", "al", "it", "as", "is", "ha", "et", "se", "ou", "of", "le", "sa", "ve", "ro",
           "ra", "ri", "hi", "ne", "me", "de", "co", "ta", "ec", "si", "ll", "so", "na", "li", "la", "el"]  # cphoc.c:30
PHOC_DIM = 604
_f = np.float32


def clean_token(token):
    """build_phoc.py:10-12."""
    token = token.lower().strip()
    return "".join(c for c in token if c in UNIGRAMS)


def phoc_of_clean(word):
    """cphoc.c:25-113 for a word that only holds [a-z0-9]."""
    phoc = np.zeros(PHOC_DIM, np.float32)
    n = len(word)
    for index in range(n):                                              # cphoc.c:33
        occ0 = _f(index) / _f(n)
        occ1 = _f(index + 1) / _f(n)
        ci = UNIGRAMS.index(word[index])
        for level in range(2, 6):                                       # cphoc.c:54
            for region in range(level):
                r0 = _f(region) / _f(level)
                r1 = _f(region + 1) / _f(level)
                o0 = max(occ0, r0)
                o1 = min(occ1, r1)
                kkk = _f(_f(o1 - o0) / _f(occ1 - occ0))
                if kkk >= _f(0.5):
                    s = sum(l for l in range(2, 6) if l < level)        # cphoc.c:66-67
                    phoc[s * 36 + region * 36 + ci] = 1
    off = 36 * 14                                                       # cphoc.c:76
    for i in range(n - 1):
        bg = word[i:i + 2]
        if bg not in BIGRAMS:
            continue
        k = BIGRAMS.index(bg)
        g0 = _f(i) / _f(n)
        g1 = _f(i + 2) / _f(n)
        for region in range(2):                                         # cphoc.c:95-106
            r0 = _f(region) / _f(2)
            r1 = _f(region + 1) / _f(2)
            o0 = max(g0, r0)
            o1 = min(g1, r1)
            if _f(_f(o1 - o0) / _f(g1 - g0)) >= _f(0.5):
                phoc[off + region * 50 + k] = 1
    return phoc


def build_phoc(token):
    """build_phoc.py:9-14."""
    with np.errstate(all="ignore"):
        return phoc_of_clean(clean_token(token))


def phoc_processor(tokens, max_length):
    """processors.py:912-928 (PAD_INDEX = 0, processors.py:205)."""
    out = np.zeros((max_length, PHOC_DIM), np.float32)
    for i, t in enumerate(tokens[:max_length]):
        out[i] = build_phoc(t)
    return out
