"""Input featurisation on the device -- the step immediately before the forward path.

The reference computes the 604-d PHOC descriptor of every OCR token on the CPU inside the DataLoader workers
(`PhocProcessor`, reference pythia/datasets/processors.py:904-928, calling the C extension
pythia/utils/phoc/src/cphoc.c through pythia/utils/phoc/build_phoc.py) and ships `context_feature_1`
([960, 604] fp32 = 2.3 MB per sample, 57 % of all input bytes) to the GPU.  Here the token *bytes* travel
(~8 B per token) and `t2s_phoc_build` (csrc/featurize.cu) writes the same fp32 rows in HBM, bit-identical to the
reference binary.  `PhocProcessor` below is the same plugin (registry key "phoc", same call contract and output
dict); `phoc_batch` builds `context_feature_1` for a whole batch in one launch.

There is no CPU fallback: without the CUDA library / a GPU these calls raise.
"""
import numpy as np
import torch

from . import lib as _lib
from .pythia_api import registry

PHOC_DIM = 604


def pack_tokens(tokens):
    """tokens (list of str) -> (uint8 bytes back to back, int32 offsets[len + 1]) as pinned-able CPU tensors.

    ASCII tokens go as they are (the kernel lower-cases and filters); a token with non-ASCII characters is
    lower-cased here with Python's own `str.lower()` (what build_phoc.py:10 applies), because Unicode case
    mapping can produce ASCII letters (e.g. U+212A KELVIN SIGN -> 'k')."""
    enc = [(t if t.isascii() else t.lower()).encode("utf-8") for t in tokens]
    offsets = np.zeros(len(enc) + 1, np.int32)
    if enc:
        np.cumsum([len(e) for e in enc], out=offsets[1:])
    data = np.frombuffer(b"".join(enc), np.uint8) if offsets[-1] else np.zeros(0, np.uint8)
    return torch.from_numpy(data.copy()), torch.from_numpy(offsets)


OCR_TOKEN_WIDTH = 64      # bytes per record of the `ocr_token_bytes` field (UTF-8, zero padded)


def pack_tokens_fixed(tokens, width=OCR_TOKEN_WIDTH):
    """tokens (list of str) -> uint8 [len(tokens), width], one zero-padded UTF-8 record per token: the per-sample
    `ocr_token_bytes` field ([O, width]; the batch collator stacks it to [B, O, width]) that `T2S.forward` accepts
    INSTEAD of `context_feature_1` -- the PHOC rows are then built on the device (`t2s_phoc_build_fixed`) in front of
    the OCR encoder.  Same lower-casing rule as `pack_tokens`.  A token longer than `width` bytes raises (its
    descriptor depends on every kept symbol): choose the width for the dataset."""
    out = np.zeros((len(tokens), width), np.uint8)
    for i, t in enumerate(tokens):
        e = (t if t.isascii() else t.lower()).encode("utf-8")
        if len(e) > width:
            raise ValueError("OCR token of %d bytes does not fit the %d-byte record: raise `width`" % (len(e), width))
        out[i, :len(e)] = np.frombuffer(e, np.uint8)
    return torch.from_numpy(out)


def phoc_from_records(records, out=None):
    """uint8 CUDA tensor [..., width] of token records -> fp32 [n, 604] PHOC rows (n = product of the leading dims)."""
    if not records.is_cuda or records.dtype != torch.uint8:
        raise _lib.T2SLibraryError("phoc_from_records needs a uint8 CUDA tensor (no CPU fallback)")
    records = records.contiguous()
    width = records.shape[-1]
    n = records.numel() // width
    if out is None:
        out = torch.empty(n, PHOC_DIM, dtype=torch.float32, device=records.device)
    _lib.get_lib().phoc_build_fixed(records.data_ptr(), width, n, out.data_ptr(), out.stride(0),
                                    torch.cuda.current_stream(records.device).cuda_stream)
    return out


def phoc_rows(tokens, rows=None, device=None, out=None):
    """PHOC rows of `tokens` -> fp32 [rows, 604] on `device` (rows >= len(tokens); extra rows are zero, the
    processor's PAD_INDEX fill).  One H2D copy of the packed bytes + one kernel launch on the current stream."""
    n = len(tokens)
    rows = n if rows is None else rows
    if rows < n:
        raise ValueError("rows (%d) < number of tokens (%d)" % (rows, n))
    if not torch.cuda.is_available():
        raise _lib.T2SLibraryError("phoc_rows needs a CUDA device: the PHOC featuriser has no CPU fallback")
    if out is None:
        device = torch.device("cuda" if device is None else device)
        if device.type != "cuda":
            raise _lib.T2SLibraryError("phoc_rows runs on a CUDA device only (no CPU fallback); got %s" % device)
        out = torch.empty(rows, PHOC_DIM, dtype=torch.float32, device=device)
    else:
        if not out.is_cuda:
            raise _lib.T2SLibraryError("phoc_rows: `out` must be a CUDA tensor (no CPU fallback)")
        assert out.dtype == torch.float32 and out.dim() == 2 and out.shape[0] == rows and out.shape[1] == PHOC_DIM
        assert out.stride(1) == 1
    L = _lib.get_lib()
    data, offsets = pack_tokens(tokens)
    with torch.cuda.device(out.device):
        d_data = data.to(out.device, non_blocking=True) if data.numel() else None
        d_off = offsets.to(out.device, non_blocking=True)
        L.phoc_build(None if d_data is None else d_data.data_ptr(), d_off.data_ptr(), n, rows, out.data_ptr(),
                     out.stride(0), torch.cuda.current_stream().cuda_stream)
    return out


def phoc_batch(token_lists, max_length, device=None):
    """`context_feature_1` of a batch: list (B) of token lists -> fp32 [B, max_length, 604], one launch.
    Each sample is truncated / zero-padded to max_length exactly like PhocProcessor."""
    flat, B = [], len(token_lists)
    for toks in token_lists:
        toks = list(toks[:max_length])
        flat.extend(toks + [""] * (max_length - len(toks)))      # the empty word has an all-zero descriptor
    return phoc_rows(flat, device=device).view(B, max_length, PHOC_DIM)


# ---------------------------------------------------------------------------------------------------------------
# Frame sampling + per-frame OCR truncate / pad / pack (reference vtextgqa/dataset.py:103-253, sample_frames :371-381)
# ---------------------------------------------------------------------------------------------------------------
def sample_frames(n_frames, num_frames):
    """1-based ids of the uniformly sampled frames of a video of n_frames frames (dataset.py:103-109, 371-381): all of them
    when there are at most num_frames, else every (n_frames // num_frames)-th starting with the first.  (The kernel
    applies the same law on the device; this is the host-side mirror for code that needs the ids, e.g. to read the
    per-frame ViT features.)"""
    if n_frames <= num_frames:
        return list(range(1, n_frames + 1))
    step = n_frames // num_frames
    return [1 + i * step for i in range(num_frames)]


def ocr_info_to_csr(ocr_info, token_processor=None, width=OCR_TOKEN_WIDTH):
    """One video's OCR info ({str(frame index from 1): [{"points": 8 numbers, "ocr": str, "ID": int}, ...]}, the .npy the
    reference loads at dataset.py:96-97) -> the flat arrays `pack_ocr_frames` takes: det_points fp32 [n, 8], det_track
    int64 [n], det_tokens uint8 [n, width], frame_ptr int32 [len(ocr_info) + 1].  A per-video, per-dataset constant:
    build it once (offline, or on first touch) and keep it next to the video.  `token_processor` = the dataset's
    `ocr_token_processor` (str -> str), applied before the bytes are packed."""
    n_info = len(ocr_info)
    pts, trk, toks, ptr = [], [], [], [0]
    for f in range(1, n_info + 1):
        for d in ocr_info[str(f)]:
            pts.append(d["points"])
            trk.append(d["ID"])
            toks.append(token_processor(d["ocr"]) if token_processor else d["ocr"])
        ptr.append(len(pts))
    return {"det_points": np.asarray(pts, np.float32).reshape(-1, 8), "det_track": np.asarray(trk, np.int64),
            "det_tokens": pack_tokens_fixed(toks, width).numpy().reshape(-1, width),
            "frame_ptr": np.asarray(ptr, np.int32)}


def pack_ocr_frames(videos, num_frames, frame_ocr_num, device=None):
    """Batch of videos -> the OCR-side Sample fields of the reference, built on the device in one launch.

    videos: list (B) of dicts with the arrays of `ocr_info_to_csr` plus "n_frames" (number of video frames) and "width",
    "height" (pixels, python numbers).  Returns CUDA tensors named like the reference's fields: ocr_bbox_coordinates
    [B, O, 4] fp32, track_id / temporal_id / ocr_mask [B, O] int64, frame_id / frame_mask [B, F] int64, frame_num /
    middel_frame_id / middel_frame_idx [B] int64, and ocr_token_bytes [B, O, width] uint8 (which `T2S.forward` takes
    instead of `context_feature_1`), O = num_frames * frame_ocr_num.  No CPU fallback."""
    if not torch.cuda.is_available():
        raise _lib.T2SLibraryError("pack_ocr_frames needs a CUDA device (no CPU fallback)")
    device = torch.device("cuda" if device is None else device)
    if device.type != "cuda":
        raise _lib.T2SLibraryError("pack_ocr_frames runs on a CUDA device only (no CPU fallback); got %s" % device)
    B, F, Of = len(videos), int(num_frames), int(frame_ocr_num)
    width = videos[0]["det_tokens"].shape[1]
    n_info = np.asarray([len(v["frame_ptr"]) - 1 for v in videos], np.int32)
    n_frames = np.asarray([v["n_frames"] for v in videos], np.int32)
    if (n_frames - 1 > n_info).any() or (n_frames < 1).any():
        raise ValueError("every video needs 1 <= n_frames <= len(ocr_info) + 1 (the reference reads frame n or n - 1)")
    det_base = np.concatenate([[0], np.cumsum([len(v["det_track"]) for v in videos])]).astype(np.int64)
    info_base = np.concatenate([[0], np.cumsum(n_info + 1)[:-1]]).astype(np.int32)       # each video keeps its own ptr[0]
    frame_ptr = np.concatenate([v["frame_ptr"].astype(np.int64) + det_base[i] for i, v in enumerate(videos)]).astype(np.int32)
    cat = lambda k, shape, dt: (np.concatenate([v[k] for v in videos]) if det_base[-1] else np.zeros(shape, dt))
    host = {"det_points": cat("det_points", (0, 8), np.float32), "det_track": cat("det_track", (0,), np.int64),
            "det_tokens": cat("det_tokens", (0, width), np.uint8), "frame_ptr": frame_ptr, "info_base": info_base,
            "n_info": n_info, "n_frames": n_frames,
            "vid_w": np.asarray([float(v["width"]) for v in videos], np.float64),
            "vid_h": np.asarray([float(v["height"]) for v in videos], np.float64)}
    O = F * Of
    with torch.cuda.device(device):
        d = {k: torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True) for k, a in host.items()}
        i64 = lambda *s: torch.empty(*s, dtype=torch.int64, device=device)
        out = {"ocr_bbox_coordinates": torch.empty(B, O, 4, dtype=torch.float32, device=device),
               "track_id": i64(B, O), "temporal_id": i64(B, O), "ocr_mask": i64(B, O), "frame_id": i64(B, F),
               "frame_mask": i64(B, F), "frame_num": i64(B), "middel_frame_id": i64(B), "middel_frame_idx": i64(B),
               "ocr_token_bytes": torch.empty(B, O, width, dtype=torch.uint8, device=device)}
        ptr = lambda t: t.data_ptr() if t.numel() else None
        _lib.get_lib().pack_ocr_frames(
            ptr(d["det_points"]), ptr(d["det_track"]), ptr(d["det_tokens"]), width, d["frame_ptr"].data_ptr(),
            d["info_base"].data_ptr(), d["n_info"].data_ptr(), d["n_frames"].data_ptr(), d["vid_w"].data_ptr(),
            d["vid_h"].data_ptr(), B, F, Of, out["ocr_bbox_coordinates"].data_ptr(), out["track_id"].data_ptr(),
            out["temporal_id"].data_ptr(), out["ocr_mask"].data_ptr(), out["frame_id"].data_ptr(),
            out["frame_mask"].data_ptr(), out["frame_num"].data_ptr(), out["middel_frame_id"].data_ptr(),
            out["middel_frame_idx"].data_ptr(), out["ocr_token_bytes"].data_ptr(),
            torch.cuda.current_stream().cuda_stream)
    return out


class PhocProcessor:
    """Drop-in for the reference's `@registry.register_processor("phoc")` (processors.py:904-928 on top of
    VocabProcessor.__call__, processors.py:237-284): `proc({"tokens": [...]})` ->
    {"text": fp32 [max_length, 604], "tokens": padded token list, "length": 0-d int64}.  "text" lives on the GPU."""
    PAD_TOKEN = "<pad>"
    PAD_INDEX = 0
    MAX_LENGTH_DEFAULT = 50

    def __init__(self, config, *args, **kwargs):
        self.config = config
        ml = getattr(config, "max_length", None) if not isinstance(config, dict) else config.get("max_length")
        self.max_length = int(ml) if ml is not None else self.MAX_LENGTH_DEFAULT
        self.device = kwargs.get("device", None)

    def _map_strings_to_indices(self, tokens):
        return phoc_rows(list(tokens[: self.max_length]), rows=self.max_length, device=self.device)

    def _pad_tokens(self, tokens):
        padded = [self.PAD_TOKEN] * self.max_length
        n = min(len(tokens), self.max_length)
        padded[:n] = tokens[:n]
        return padded, torch.tensor(n, dtype=torch.long)

    def __call__(self, item):
        if not isinstance(item, dict):
            raise TypeError("Argument passed to the processor must be a dict with either 'text' or 'tokens' as keys")
        if "tokens" not in item:
            raise AssertionError("A dict with 'tokens' must be passed to the PHOC processor")
        tokens = item["tokens"]
        indices = self._map_strings_to_indices(tokens)
        tokens, length = self._pad_tokens(tokens)
        return {"text": indices, "tokens": tokens, "length": length}


def _register():
    """Take over the "phoc" processor key.  The real registry asserts a BaseProcessor subclass
    (registry.py:200-214), so when pythia's processors module is importable the class is mixed with it."""
    cls = PhocProcessor
    try:
        from pythia.datasets.processors import BaseProcessor  # type: ignore
        cls = type("PhocProcessor", (PhocProcessor, BaseProcessor), {})
    except Exception:
        pass
    registry.mapping.setdefault("processor_name_mapping", {})["phoc"] = cls
    return cls


_register()
