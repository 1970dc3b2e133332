#!/usr/bin/env python
"""Golden vectors for the frame-sampling + OCR pad / pack step, produced by the REFERENCE's own code.

The reference does this inside `VTEXTGQADataset.add_sample_details` (pythia/datasets/videoqa/vtextgqa/dataset.py:83-287)
with `sample_frames` (dataset.py:371-381) and `CopyProcessor` (pythia/datasets/processors.py:932-944).  Importing that
module pulls in the whole dataset zoo (and third-party packages this image lacks), and the method reads the OCR .npy
files, the frame directory and the ViT feature files of the authors' machine.  So this script takes the SOURCE TEXT of
those three definitions from /root/reference with `ast` (nothing is copied into the repository), executes it unchanged
in a namespace whose file-system calls (`np.load`, `glob.glob`) answer from synthetic in-memory videos and whose
string processors are identities, and stores the arrays the method produced next to the inputs it was given:

    tests/golden/ocr_pack_golden.npz      (written by this script; read by tests/test_featurize.py)

Run here (needs /root/reference):   python tests/golden/make_pack_golden.py
"""
import ast
import os
import random
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/pythia"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ocr_pack_golden.npz")
WIDTH = 64


def source_of(path, name, cls=None):
    """Source text of a top-level function / of a method or class of the reference file."""
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if cls is None and isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name == name:
            return ast.get_source_segment(src, node)
        if cls is not None and isinstance(node, ast.ClassDef) and node.name == cls:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == name:
                    import textwrap
                    return textwrap.dedent(ast.get_source_segment(src, sub, padded=True))
    raise KeyError(name)


class Sample(dict):
    """Attribute bag standing in for pythia.common.sample.Sample (only attribute assignment is used on this path)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def make_video(rng, n_frames, n_info, of_max, empty_ratio=0.2):
    """Synthetic OCR info in the layout the reference reads: {str(frame index from 1): [{points, ocr, ID}, ...]}."""
    words = ["exit", "STOP", "coca-cola", "24h", "Main St.", "a", "open", "P", "No.7", "café", "2nd", "sale!", "x" * 40]
    info = {}
    for f in range(1, n_info + 1):
        k = 0 if rng.random() < empty_ratio else rng.randint(1, of_max)
        dets = []
        for _ in range(k):
            x, y = rng.uniform(0, 1200), rng.uniform(0, 700)
            w, h = rng.uniform(5, 200), rng.uniform(5, 80)
            jit = lambda: rng.uniform(-3, 3)
            pts = [x + jit(), y + jit(), x + w + jit(), y + jit(), x + w + jit(), y + h + jit(), x + jit(), y + h + jit()]
            if rng.random() < 0.5:
                pts = [int(p) for p in pts]           # the OCR files hold ints for most detectors
            dets.append({"points": pts, "ocr": rng.choice(words), "ID": rng.randint(1, 400)})
        info[str(f)] = dets
    return info


def encode(tokens):
    """list of str -> uint8 [n, WIDTH], zero-padded UTF-8 records (no pickled objects in the fixture)."""
    a = np.zeros((len(tokens), WIDTH), np.uint8)
    for i, t in enumerate(tokens):
        e = t.encode("utf-8")
        a[i, :len(e)] = np.frombuffer(e, np.uint8)
    return a


def main():
    ds_path = os.path.join(REF, "datasets/videoqa/vtextgqa/dataset.py")
    method_src = source_of(ds_path, "add_sample_details", cls="VTEXTGQADataset")
    sample_frames_src = source_of(ds_path, "sample_frames")
    copy_src = source_of(os.path.join(REF, "datasets/processors.py"), "CopyProcessor")

    state = {}

    def fake_load(path, allow_pickle=False):
        if "ocr_dir" in path:                       # the per-video OCR info (np.load(...).item())
            class Box:
                def item(self_inner):
                    return state["info"]
            return Box()
        return np.zeros((1, 1024), np.float32)      # a ViT feature file

    fake_np = types.SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
    fake_np.load = fake_load
    fake_glob = types.SimpleNamespace(glob=lambda pattern: ["%d.jpg" % i for i in range(state["n_frames"])])
    ns = {"np": fake_np, "torch": torch, "os": os, "glob": fake_glob, "Sample": Sample, "F": F,
          "enc_obj2bytes": lambda x: torch.zeros(1, dtype=torch.uint8), "BaseProcessor": object}
    exec(sample_frames_src, ns)
    exec(copy_src, ns)
    exec(method_src, ns)
    add_sample_details = ns["add_sample_details"]

    cases = [  # (video frames, len(ocr_info), F, Of, max detections per frame)
        (40, 40, 16, 5, 9),       # more frames than F: step 2, frames with more detections than Of
        (16, 16, 16, 5, 4),       # exactly F
        (9, 9, 16, 5, 6),         # fewer frames than F: zero-padded frame slots
        (50, 49, 16, 5, 5),       # OCR info one frame short of the frame directory: the `frame_idx - 1` branch
        (130, 130, 64, 15, 20),   # the headline geometry
        (7, 7, 64, 15, 3),
    ]
    rng = random.Random(20261017)
    out = {"n_cases": np.int64(len(cases)), "width": np.int64(WIDTH)}
    for ci, (n_frames, n_info, Fn, Of, of_max) in enumerate(cases):
        info = make_video(rng, n_frames, n_info, of_max)
        state.update(info=info, n_frames=n_frames)
        O = Fn * Of
        self = types.SimpleNamespace(
            num_frames=Fn, frame_ocr_num=Of, ocr_info_dir=["ocr_dir"],
            text_processor=lambda d: {"token_inds": torch.zeros(20, dtype=torch.long), "token_num": torch.tensor(3)},
            copy_processor=ns["CopyProcessor"](types.SimpleNamespace(max_length=O)),
            ocr_token_processor=lambda d: {"text": d["text"]},
            context_processor=lambda d: {"text": torch.zeros(1), "tokens": list(d["tokens"]), "length": torch.tensor(len(d["tokens"]))},
            phoc_processor=lambda d: {"text": torch.zeros(1), "length": torch.tensor(len(d["tokens"]))})
        vw, vh = rng.choice([(1280, 720), (1920, 1080), (640, 360), (1000.5, 562.75)])
        sample_info = {"question": "what", "video_id": "v%d" % ci, "video_width": vw, "video_height": vh}
        s = add_sample_details(self, sample_info, Sample())
        tokens = list(sample_info["ocr_tokens"])
        # ---- inputs in the layout of t2s_pack_ocr_frames: detections of all info frames back to back (CSR)
        ptr, pts, trk, toks = [0], [], [], []
        for f in range(1, n_info + 1):
            for d in info[str(f)]:
                pts.append(d["points"])
                trk.append(d["ID"])
                toks.append(d["ocr"])
            ptr.append(len(pts))
        p = "c%d_" % ci
        out[p + "geom"] = np.asarray([n_frames, n_info, Fn, Of], np.int64)
        out[p + "size"] = np.asarray([vw, vh], np.float64)
        out[p + "frame_ptr"] = np.asarray(ptr, np.int32)
        out[p + "det_points"] = np.asarray(pts, np.float32).reshape(-1, 8)
        out[p + "det_track"] = np.asarray(trk, np.int64)
        out[p + "det_tokens"] = encode(toks)
        # ---- what the reference produced
        out[p + "ocr_bbox_coordinates"] = s.ocr_bbox_coordinates.numpy()
        for k in ("track_id", "temporal_id", "ocr_mask", "frame_id", "frame_mask", "frame_num", "middel_frame_id",
                  "middel_frame_idx"):
            out[p + k] = s[k].numpy()
        out[p + "ocr_tokens"] = encode(tokens)      # len(idxs) * Of records, the literal "<pad>" on padded slots
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
