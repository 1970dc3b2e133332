"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's frame sampling + per-frame OCR pad / pack
("pad/pack ..., frame sampling -> feeds K3 directly").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(vitxt_gqa_b200/) never does.

Follows, line by line:
  * pythia/datasets/videoqa/vtextgqa/dataset.py:371-381  -- sample_frames(frames, sample_len, 'uniform'): every frame
    when there are at most sample_len, else frames[i * (len // sample_len)]
  * dataset.py:103-109    -- frame ids 1 .. len(frame_paths)
  * dataset.py:116-158    -- per sampled frame: ocr_info[str(frame_idx)] (or str(frame_idx - 1) when the OCR info is one
    frame short), box = (min(x0, x6), min(y1, y3), max(x2, x4), max(y5, y7)) of the eight quadrilateral coordinates,
    track id, temporal id = frame index; truncated / padded to frame_ocr_num slots with "<pad>", a zero box, track 0,
    the SAME frame index, mask 0
  * dataset.py:166-195    -- middel_frame_id = the LAST sampled frame (the third assignment wins), middel_frame_idx =
    len(frames) // 2 + 1 when that id is >= num_frames, else the id itself
  * dataset.py:199-243    -- frame ids / masks zero-padded to num_frames; boxes as float32, multiplied by the python
    floats [1/width, 1/height, 1/width, 1/height] (a float64 product) and cast back to float32; CopyProcessor
    (processors.py:932-944) zero-pads to num_frames * frame_ocr_num rows; track / temporal / mask vectors zero-padded
  * dataset.py:246-253    -- the token list has len(sampled frames) * frame_ocr_num entries; the text processors pad the
    rest (PHOC: zero rows, processors.py:904-928) -- here: empty records

Pinning: tests/golden/make_pack_golden.py executes the reference's own source text of these functions over synthetic
in-memory videos and commits tests/golden/ocr_pack_golden.npz; tests/test_oracle_cpu.py holds this file to it bit for bit.
"""
import numpy as np

PAD_TOKEN = b"<pad>"


def sample_frames(n_frames, num_frames):
    """1-based ids of the sampled frames (dataset.py:103-109, 371-381)."""
    frames = list(range(1, n_frames + 1))
    if len(frames) <= num_frames:
        return frames
    step = len(frames) // num_frames
    return [frames[i * step] for i in range(num_frames)]


def pack_ocr_frames(det_points, det_track, det_tokens, frame_ptr, n_info, n_frames, width_px, height_px, num_frames,
                    frame_ocr_num):
    """One video.  det_* hold the detections of OCR-info frame 1, 2, ... back to back, frame_ptr[j] .. frame_ptr[j + 1]
    being those of frame j + 1; det_tokens is uint8 [n, W] (zero-padded records).  Returns a dict of numpy arrays named as
    the reference names the sample fields, plus `ocr_token_bytes` uint8 [num_frames * frame_ocr_num, W]."""
    F, Of = num_frames, frame_ocr_num
    O, W = F * Of, det_tokens.shape[1]
    idxs = sample_frames(n_frames, F)
    bbox = np.zeros((O, 4), np.float32)
    track = np.zeros(O, np.int64)
    temporal = np.zeros(O, np.int64)
    mask = np.zeros(O, np.int64)
    tokens = np.zeros((O, W), np.uint8)
    frame_id = np.zeros(F, np.int64)
    frame_mask = np.zeros(F, np.int64)
    scale = np.array([1.0 / width_px, 1.0 / height_px, 1.0 / width_px, 1.0 / height_px], np.float64)
    pad = np.zeros(W, np.uint8)
    pad[:len(PAD_TOKEN)] = np.frombuffer(PAD_TOKEN, np.uint8)
    for i, frame_idx in enumerate(idxs):
        j = frame_idx if n_info >= frame_idx else frame_idx - 1                       # dataset.py:121-124
        lo, hi = int(frame_ptr[j - 1]), int(frame_ptr[j])
        k = min(hi - lo, Of)
        for o in range(Of):
            s = i * Of + o
            temporal[s] = frame_idx
            if o < k:
                p = det_points[lo + o].astype(np.float32)
                box = np.array([min(p[0], p[6]), min(p[1], p[3]), max(p[2], p[4]), max(p[5], p[7])], np.float32)
                bbox[s] = (box.astype(np.float64) * scale).astype(np.float32)        # dataset.py:208-212
                track[s] = det_track[lo + o]
                mask[s] = 1
                tokens[s] = det_tokens[lo + o]
            else:
                tokens[s] = pad
        frame_id[i] = frame_idx
        frame_mask[i] = 1
    last = idxs[-1]
    mid_idx = len(idxs) // 2 + 1 if last >= F else last                                   # dataset.py:186-190
    return {"ocr_bbox_coordinates": bbox, "track_id": track, "temporal_id": temporal, "ocr_mask": mask,
            "frame_id": frame_id, "frame_mask": frame_mask, "frame_num": np.int64(len(idxs)),
            "middel_frame_id": np.array([last], np.int64), "middel_frame_idx": np.array([mid_idx], np.int64),
            "ocr_token_bytes": tokens}
