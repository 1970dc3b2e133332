"""Training step of T2S on the B200 kernels: teacher-forced forward that keeps what the backward needs,
hand-scheduled backward, flat-buffer gradient all-reduce, fused clip + Adam.

Reference: one training step of pythia/trainers/base_trainer.py:226-270 -- `model(prepared_batch)` in training
mode (models/t2s.py:288-314: three teacher-forced MMT passes ref / pos / neg), `loss.backward()`,
`clip_grad_norm_(model.parameters(), 0.25)` (utils/general.py:32-40), `Adam.step()`; under
DistributedDataParallel the gradients are all-reduced (mean) inside backward (base_trainer.py:134-137).

How it plugs in: `T2S.forward(sample_list)` in training mode with autograd enabled routes through
`_T2STrainFn`, a `torch.autograd.Function` whose forward enqueues the kernel schedule below and whose backward
enqueues the backward schedule and returns the parameter gradients, so the reference's trainer
(`loss.backward(); clip; optimizer.step()`) works unchanged.  `TrainEngine.step()` is the fused alternative to
clip_grad_norm_ + torch.optim.Adam over the same flat buffers, and `TrainEngine.all_reduce()` the NCCL gradient
all-reduce of one flat fp32 buffer (live parameters only: the 19.5 M dead parameters of SURVEY Q18 never get a
gradient and are left out, which is what DDP's find_unused_parameters=True arranges in the reference).

Arithmetic: forward exactly as the eval path (fp32-class bf16x3 grounding chain, bf16 answer transformer) so the
grounding indices match; backward in bf16 with fp32 accumulation (activation gradients bf16, parameter gradients
fp32).  Dropout: the reference trains with p = 0.1 at BertEmbeddings, every BertSelfOutput / BertOutput, the attention
probabilities, obj_drop / ocr_drop and PrevPredEmbeddings.emb_dropout (models/t2s.py:95,118,688,720 and the BertConfig
defaults behind t2s.py:25-28); the engine reads the same probabilities from the model config and applies them with
counter-based masks that the backward recomputes from (seed, site) -- nothing is stored (include/t2s_b200.h,
"Dropout of the training step").  Parity with autograd is defined at p = 0 (`eng.set_dropout(False)`, or zero
probabilities in the config); at p > 0 the masks are checked statistically and forward against backward.

Schedule of the backward (per variant v in ref, pos, neg; then the shared front):
  loss -> dscores (bf16 copy) -> classifier / pointer-net dgrad + wgrad -> for each answer-transformer layer, last to
  first, decoder rows and encoder rows: LN2 bwd -> FFN-down dgrad (x GELU') + wgrad -> FFN-up dgrad (+ residual)
  + wgrad -> LN1 bwd -> attention-out dgrad + wgrad -> attention bwd (encoder + decoder rows jointly, prefix-LM
  mask) -> q|k|v dgrad (+ residual) + wgrad -> PrevPredEmbeddings bwd; encoder-input gradients of the three
  variants summed into dJ1 -> QTV (tanh residual) -> obj / OCR encoders, TextBert -> embeddings.
"""
import os

import torch

from . import lib as _lib

LN_EPS_BERT = 1e-12
LN_EPS_EMBED = 1e-5
H = 768


def _ptr(t):
    return None if t is None else t.data_ptr()


def _ru(x, m):
    return (x + m - 1) // m * m


def lr_lambda_update(i_iter, cfg):
    """The reference's learning-rate multiplier (pythia/utils/general.py:20-30): linear warm-up from `warmup_factor`
    to 1 over `warmup_iterations`, then `lr_ratio` ** (number of `lr_steps` passed).  `TrainEngine.step(lr=base_lr *
    lr_lambda_update(i, cfg), ...)` is what `LambdaLR` does for the reference's optimizer."""
    from bisect import bisect
    tp = cfg["training_parameters"]
    if tp["use_warmup"] is True and i_iter <= tp["warmup_iterations"]:
        alpha = float(i_iter) / float(tp["warmup_iterations"])
        return tp["warmup_factor"] * (1.0 - alpha) + alpha
    return pow(tp["lr_ratio"], bisect(tp["lr_steps"], i_iter))


class _T2STrainFn(torch.autograd.Function):
    """forward(engine, inputs, *live_params) -> one score tensor per entry of `engine.out_variants`
    (T2S: ref, pos, neg; the single-variant models: one)."""

    @staticmethod
    def forward(ctx, engine, inp, *params):
        ctx.engine = engine
        out = tuple(engine.forward(inp))
        ctx.fwd_gen = engine.fwd_gen        # the engine keeps ONE set of saved activations: this forward's
        return out

    @staticmethod
    def backward(ctx, *dscores):
        eng = ctx.engine
        if eng.saved is None or ctx.fwd_gen != eng.fwd_gen:
            raise RuntimeError(
                "T2S training backward: the saved activations belong to a different forward (the engine keeps one "
                "forward's activations; run forward -> backward pairs, not two training forwards before one backward)")
        eng.backward(dict(zip(eng.out_variants, dscores)))
        # autograd owns what is returned here (AccumulateGrad steals it as p.grad or adds it to an existing p.grad), so
        # it must not alias the engine's flat buffer, which the next backward zeroes and refills: hand out views of ONE
        # copy of the live range (351 MB, ~0.1 ms).  TrainEngine.all_reduce() / .step() keep using the flat buffer.
        return (None, None) + tuple(eng.grad_copies())


class TrainEngine:
    """Flat parameter / gradient / Adam-state buffers of a T2S model + the training kernel schedules."""

    def __init__(self, model):
        self.model = model
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise _lib.T2SLibraryError("the training step runs only on a CUDA device (sm_100a); there is no CPU fallback")
        if model.grounding_precision != "bf16x3":
            raise NotImplementedError("training uses the bf16x3 grounding chain")
        # T2S and its ablations: three variants behind the QTV layers; M4C / T5-ViteVQA / GT-box: ONE answer-transformer
        # variant straight on the encoders' output (no QTV, no grounding gradient), driven through the model's _sv_* hooks
        self.is_t2s = model.MODEL.startswith("t2s")
        if self.is_t2s:
            self.variants = ("pos", "ref", "neg")
            self.out_variants = ("ref", "pos", "neg")
        else:
            if getattr(model, "TRAIN_VARIANT", None) is None:
                raise NotImplementedError("no training schedule for model '%s'" % model.MODEL)
            self.variants = self.out_variants = (model.TRAIN_VARIANT,)
        self.DEAD_PREFIXES = tuple(getattr(model, "TRAIN_DEAD", self.DEAD_PREFIXES))
        self._flatten()
        self.step_count = 0
        self._ws = {}
        self._wt = None
        self._wt_key = None
        self.saved = None
        self.fwd_gen = 0
        self._read_dropout_config()
        self.grad_scale = 1.0      # set by all_reduce(): the factor that turns the summed gradients into their mean
        # gradient all-reduce overlapped with the backward (reference: DDP's bucketed all-reduce inside autograd,
        # base_trainer.py:134-137): as soon as a contiguous range of the flat gradient buffer is final, its NCCL
        # all-reduce is enqueued on a side stream behind an event of the compute stream.  Built, tested (gradient parity
        # at N = 2 and 8) and measured on one 8 x B200 NVSwitch box -- where it LOSES to one flat all-reduce after the
        # backward: NCCL's CTAs compete with the persistent tcgen05 grids of the backward for SMs, so the "hidden"
        # transfers slow the backward by more than the < 1 ms a 351 MB NVLS all-reduce costs on its own (batch 48 / GPU:
        # 55.5 / 55.1 ms per step with 64 / 128 MB buckets, 54.8 ms with the flat call; batch 6 / GPU: 17.35 vs 16.94 ms).
        # Default therefore: the flat call (T2S_B200_OVERLAP_ALLREDUCE=1 turns the overlap on, e.g. for PCIe / multi-node).
        self.overlap_allreduce = os.environ.get("T2S_B200_OVERLAP_ALLREDUCE", "0") not in ("0", "False")
        # final ranges are collected until this many bytes are pending, then go out as one group of NCCL calls: on
        # NVSwitch a 351 MB all-reduce alone takes < 1 ms, while many small all-reduces under the backward are SM-starved by
        # the persistent GEMM grids and slow the backward down more than they hide (measured at N = 8)
        self.bucket_bytes = int(float(os.environ.get("T2S_B200_ALLREDUCE_BUCKET_MB", "64")) * (1 << 20))
        self._ready_ranges = []
        self.reduce_in_backward = True     # False: leave the local gradients alone (a caller that reduces by itself)
        self._comm_stream = None
        self._buckets_pending = []     # [(lo, hi)] of this backward, in launch order
        self._reduced_upto = None      # None: nothing reduced by the current backward
        self.comm_events = []          # measurement aid: [(bytes, start_event, end_event)] of the last backward
        self.time_comm = False

    # ------------------------------------------------------------------ dropout
    def _read_dropout_config(self):
        """Dropout probabilities exactly where the reference reads them: `config.obj.dropout_prob`,
        `config.ocr.dropout_prob` (t2s.py:95,118) and the BertConfig built from `config.text_bert` / `translayers` /
        `mmt` (t2s.py:25-28,46), whose `hidden_dropout_prob` / `attention_probs_dropout_prob` default to 0.1."""
        cfg = self.model.config

        def bert(key):
            node = cfg.get(key, None) if hasattr(cfg, "get") else None
            node = dict(node) if node is not None else {}
            return float(node.get("hidden_dropout_prob", 0.1)), float(node.get("attention_probs_dropout_prob", 0.1))

        def prob(key):
            node = cfg.get(key, None) if hasattr(cfg, "get") else None
            return float(dict(node).get("dropout_prob", 0.0)) if node is not None else 0.0

        self.drop_cfg = dict(obj=prob("obj"), ocr=prob("ocr"), text=bert("text_bert"), qtv=bert("translayers"),
                             mmt=bert("mmt"))
        self.dropout_enabled = True
        self._site_ids = {}
        self.base_seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF

    def set_dropout(self, enabled, seed=None):
        """Switch the training-step dropout on / off (off = the p = 0 step the autograd parity tests are defined
        at); `seed` re-bases the mask sequence."""
        self.dropout_enabled = bool(enabled)
        if seed is not None:
            self.base_seed = int(seed) & 0xFFFFFFFFFFFFFFFF

    @property
    def dropout_p(self):
        """The largest dropout probability in effect (0.0 when dropout is off) -- what bench.py reports."""
        if not (self.dropout_enabled and getattr(self.model, "train_dropout", True)):
            return 0.0
        d = self.drop_cfg
        return max(d["obj"], d["ocr"], *d["text"], *(d["qtv"] if self.is_t2s else (0.0,)), *d["mmt"])

    def _p(self, key, which=None):
        if not (self.dropout_enabled and getattr(self.model, "train_dropout", True)):
            return 0.0
        v = self.drop_cfg[key]
        return v if which is None else v[which]

    def _site(self, name):
        """Small integer id of a dropout site (stable across steps; forward and backward name sites identically)."""
        i = self._site_ids.get(name)
        if i is None:
            i = self._site_ids[name] = len(self._site_ids) + 1
            if i >= 4096:
                raise RuntimeError("too many dropout sites")
        return i

    def _step_seed(self):
        r = 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                r = dist.get_rank()
        except Exception:
            pass
        return (self.base_seed + self.fwd_gen * 0x9E3779B97F4A7C15 + r * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF

    # ------------------------------------------------------------------ flat buffers
    DEAD_PREFIXES = ("Grounding_Module.", "linear_obj_frame_to_mmt_in.", "obj_frame_layer_norm.")

    def _flatten(self):
        """Re-point every parameter at a view of one flat fp32 buffer, ordered (a) by optimizer group (reference
        get_optimizer_parameters, t2s.py:356-376: default, text_bert x lr_scale_text_bert, mmt x lr_scale_mmt),
        dead parameters last, and (b) with query/key/value weights (and biases) adjacent, so the fused q|k|v
        weight gradient is one [3H, H] view."""
        m = self.model
        named = dict(m.named_parameters())
        text = [n for n in named if n.startswith("text_bert.")]
        mmt = [n for n in named if n.startswith("mmt.")]
        dead = [n for n in named if n.startswith(self.DEAD_PREFIXES)]
        taken = set(text) | set(mmt) | set(dead)
        default = [n for n in named if n not in taken]

        def order(names):
            def key(n):
                # q, k, v of one attention block adjacent and in that order; weights before biases
                for i, tag in enumerate(("query", "key", "value")):
                    if ".attention.self." + tag + "." in n:
                        base = n.split(".attention.self.")[0]
                        return (base + ".attention.self", 0 if n.endswith("weight") else 1, i)
                return (n, 2, 0)
            return sorted(names, key=key)

        finetune_text = any(f["module"] is m.text_bert for f in m.finetune_modules)
        groups = [("default", order(default + ([] if finetune_text else text))),
                  ("text_bert", order(text) if finetune_text else []),
                  ("mmt", order(mmt)), ("dead", order(dead))]
        total = sum(named[n].numel() for _, ns in groups for n in ns)
        # every segment starts 16-byte aligned (TMA reduce-add / float4 access): pad each tensor to 4 floats
        sizes = {n: _ru(named[n].numel(), 4) for n in named}
        total = sum(sizes.values())
        flat_p = torch.zeros(total, device=self.dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=self.dev, dtype=torch.float32)
        self.adam_m = torch.zeros(total, device=self.dev, dtype=torch.float32)
        self.adam_v = torch.zeros(total, device=self.dev, dtype=torch.float32)
        self.offsets, self.group_ranges = {}, {}
        off = 0
        for gname, names in groups:
            start = off
            for n in names:
                p = named[n]
                view = flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                self.offsets[n] = off
                off += sizes[n]
            self.group_ranges[gname] = (start, off)
        self.flat_param = flat_p
        self.live_end = self.group_ranges["mmt"][1]
        self.named = named
        self.live_names = [n for g, ns in groups[:3] for n in ns]
        self.live_params = [named[n] for n in self.live_names]

    def grad(self, name):
        p = self.named[name]
        o = self.offsets[name]
        return self.flat_grad[o:o + p.numel()].view_as(p)

    def _g(self, name):
        return self.flat_grad.data_ptr() + 4 * self.offsets[name]

    def grad_copies(self):
        """Per-parameter views of one fresh copy of the live gradient range, in `live_names` order."""
        flat = self.flat_grad[:self.live_end].clone()
        if self._reduced_upto == self.live_end and self.grad_scale != 1.0:
            flat.mul_(self.grad_scale)        # DDP leaves the MEAN over ranks in p.grad
        return [flat[self.offsets[n]:self.offsets[n] + self.named[n].numel()].view_as(self.named[n])
                for n in self.live_names]

    # ------------------------------------------------------------------ transposed bf16 weights for the dgrad GEMMs
    def _wt_pack(self):
        m = self.model
        key = m._weights_key()
        if self._wt is not None and self._wt_key == key:
            return self._wt
        bf = lambda w: w.detach().to(torch.bfloat16)

        def layer(l):
            a = l.attention
            wqkv = torch.cat([a.self.query.weight, a.self.key.weight, a.self.value.weight], 0)
            return dict(wqkvT=bf(wqkv).t().contiguous(), woT=bf(a.output.dense.weight).t().contiguous(),
                        wiT=bf(l.intermediate.dense.weight).t().contiguous(),
                        wo2T=bf(l.output.dense.weight).t().contiguous())

        W = dict(text=[layer(l) for l in m.text_bert.encoder.layer],
                 qtv=[layer(l) for l in m.TransLayer.encoder.layer] if self.is_t2s else [],
                 mmt=[layer(l) for l in m.mmt.encoder.layer])
        V = m.classifier.module.weight.shape[0]
        clsT = torch.zeros(H, _ru(V, 8), device=self.dev, dtype=torch.bfloat16)
        clsT[:, :V] = bf(m.classifier.module.weight).t()
        W["clsT"] = clsT
        W["ptr_qT"] = bf(m.ocr_ptr_net.query.weight).t().contiguous()
        W["ptr_kT"] = bf(m.ocr_ptr_net.key.weight).t().contiguous()
        P = m._packed
        for which, lin in (("obj", m.linear_obj_feat_to_mmt_in), ("ocr", m.linear_ocr_feat_to_mmt_in)):
            kp = P["k_%s_pad" % which]
            t = torch.zeros(kp, H, device=self.dev, dtype=torch.bfloat16)
            t[:lin.weight.shape[1]] = bf(lin.weight).t()
            W[which + "T"] = t
        self._wt, self._wt_key = W, key
        return W

    # ------------------------------------------------------------------ workspaces
    def _train_ws(self, B, Lt, F, O, T, V, kp_obj, kp_ocr):
        """`F` = rows of the object part of the joint sequence (frames; 1 for M4C)."""
        key = (B, Lt, F, O, T, V)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        dev = self.dev
        Le = Lt + F + O
        Me, Md, Mt = B * Le, B * T, B * Lt
        f32 = dict(device=dev, dtype=torch.float32)
        b16 = dict(device=dev, dtype=torch.bfloat16)
        m = self.model
        n_text, n_mmt = len(m.text_bert.encoder.layer), len(m.mmt.encoder.layer)
        n_qtv = len(m.TransLayer.encoder.layer) if self.is_t2s else 0
        Np = _ru(V + O, 8)

        def x3_layer(M):      # what one grounding-chain layer keeps (bf16 hi|lo operands double as bf16 activations)
            return dict(stats=torch.empty(M * 12 * 2, **f32),      # {log2-sum-exp, dO . O} per (sample, head, row)
                        xs=torch.empty(M, 2 * H, **b16), qkvs=torch.empty(M, 6 * H, **b16), ctxs=torch.empty(M, 2 * H, **b16),
                        h1=torch.empty(M, H, **f32), x1=torch.empty(M, H, **f32), x1s=torch.empty(M, 2 * H, **b16),
                        u=torch.empty(M, 4 * H, **f32), inters=torch.empty(M, 8 * H, **b16), h2=torch.empty(M, H, **f32),
                        out=torch.empty(M, H, **f32))

        def bf_layer(M, need_qkv=True):
            d = dict(ctx=torch.empty(M, H, **b16), h1=torch.empty(M, H, **b16), x1=torch.empty(M, H, **b16),
                     u=torch.empty(M, 4 * H, **b16), inter=torch.empty(M, 4 * H, **b16), h2=torch.empty(M, H, **b16),
                     out=torch.empty(M, H, **b16))
            if need_qkv:
                d["qkv"] = torch.empty(M, 3 * H, **b16)
            return d

        variants = self.variants
        ws = dict(
            text=[x3_layer(Mt) for _ in range(n_text)], qtv=[x3_layer(Me) for _ in range(n_qtv)],
            xt=torch.empty(Mt, H, **f32),
            enc={v: [dict(bf_layer(Me, need_qkv=(li > 0)), stats=torch.empty((Me + Md) * 12 * 2, **f32))
                     for li in range(n_mmt)] for v in variants},
            dec={v: [bf_layer(Md) for _ in range(n_mmt)] for v in variants},
            xd={v: torch.empty(Md, H, **b16) for v in variants},
            qd={v: torch.empty(Md, H, **b16) for v in variants},
            keyp={v: torch.empty(Me, H, **b16) for v in variants},
            # backward scratch (encoder-row sized; decoder rows use the *_d copies)
            dy=torch.empty(Me, H, **b16), dy2=torch.empty(Me, H, **b16), dh2=torch.empty(Me, H, **b16),
            du=torch.empty(Me, 4 * H, **b16), dx1=torch.empty(Me, H, **b16), dh1=torch.empty(Me, H, **b16),
            dctx=torch.empty(Me, H, **b16), dqkv=torch.empty(Me, 3 * H, **b16),
            dy_d=torch.empty(Md, H, **b16), dy2_d=torch.empty(Md, H, **b16), dh2_d=torch.empty(Md, H, **b16),
            du_d=torch.empty(Md, 4 * H, **b16), dx1_d=torch.empty(Md, H, **b16), dh1_d=torch.empty(Md, H, **b16),
            dctx_d=torch.empty(Md, H, **b16), dqkv_d=torch.empty(Md, 3 * H, **b16),
            # dropout: gradients of the Linear outputs behind the hidden dropouts (masked copies of dh2 / dh1)
            dh2m=torch.empty(Me, H, **b16), dh1m=torch.empty(Me, H, **b16),
            dh2m_d=torch.empty(Md, H, **b16), dh1m_d=torch.empty(Md, H, **b16),
            dkeyp=torch.zeros(Me, H, **b16), dq=torch.empty(Md, H, **b16),
            dS16=torch.zeros(Md, Np, **b16), dJ=torch.empty(Me, H, **f32),
            dh_obj=torch.empty(B * F, H, **b16), dh_ocr=torch.empty(B * O, H, **b16), dc_ocr=torch.empty(B * O, H, **f32),
            dw_obj=torch.empty(H, kp_obj, **f32), dw_ocr=torch.empty(H, kp_ocr, **f32),
            d_id_obj=torch.empty(B * F, 52, **f32), d_id_ocr=torch.empty(B * O, 100, **f32),
            attn_ws=torch.empty(int(_lib.get_lib().attn_bwd_workspace_bytes(B, Le, T, 12)), device=dev, dtype=torch.uint8),
            Np=Np,
        )
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ forward building blocks
    def _x3_layer_fwd(self, L, lw, x, sv, M, rows_L, keys, nk, key_stride, st, out, tanh_base=None, out16=None,
                      remap=(0, 0, 0), next_xs=None, drop=None):
        """One grounding-chain BERT layer (as model._layer_f32, bf16x3), keeping its intermediates in `sv`.
        sv["xs"] already holds the bf16 hi|lo split of the input x.  drop = (p_hidden, p_attn, seed, site name)."""
        B = M // rows_L
        F32, RES, SPLIT = _lib.GEMM_OUT_F32, _lib.GEMM_RES_F32, _lib.GEMM_OUT_SPLIT
        p_h, p_a, seed, sname = drop if drop is not None else (0.0, 0.0, 0, "")
        L.gemm_bf16x3(_ptr(sv["xs"]), 2 * H, _ptr(lw["wqkv"]), 2 * H, _ptr(lw["bqkv"]), None, 0, _ptr(sv["qkvs"]), 6 * H,
                      M, 3 * H, H, SPLIT, 0, st)
        # the training form of the attention kernel: dropout on the probabilities (p_a may be 0) and the rows'
        # log2-sum-exp saved for the backward
        L.attn_tc_dropout(_ptr(sv["qkvs"]), 6 * H, 3 * H, B, rows_L, H, 12, _ptr(keys), _ptr(nk), key_stride,
                          _ptr(sv["ctxs"]), 2 * H, p_a, seed, self._site(sname + ".attn"), _ptr(sv["stats"]), rows_L, st)
        if p_h > 0:
            # BertSelfOutput / BertOutput with dropout: the GEMM leaves Linear(x) + bias, the LayerNorm kernel does
            # LN(dropout(.) + input) and writes the pre-LayerNorm sum back for the backward
            L.gemm_bf16x3(_ptr(sv["ctxs"]), 2 * H, _ptr(lw["wo"]), 2 * H, _ptr(lw["bo"]), None, 0, _ptr(sv["h1"]), H,
                          M, H, H, F32, 0, st)
            L.add_ln_dropout(_ptr(sv["h1"]), 0, H, _ptr(x), 0, H, _ptr(lw["ln1g"]), _ptr(lw["ln1b"]), LN_EPS_BERT, M, H,
                             None, 0, _ptr(sv["x1"]), H, _ptr(sv["x1s"]), 2 * H, 1, 0, 0, 0, _ptr(sv["h1"]), p_h, seed,
                             self._site(sname + ".h1"), st)
        else:
            L.gemm_bf16x3(_ptr(sv["ctxs"]), 2 * H, _ptr(lw["wo"]), 2 * H, _ptr(lw["bo"]), _ptr(x), H, _ptr(sv["h1"]), H,
                          M, H, H, F32 | RES, 0, st)
            L.add_ln_split(_ptr(sv["h1"]), 0, H, None, 0, 0, _ptr(lw["ln1g"]), _ptr(lw["ln1b"]), LN_EPS_BERT, M, H, None, 0,
                           _ptr(sv["x1"]), H, _ptr(sv["x1s"]), 2 * H, 0, 0, 0, st)
        L.gemm_bf16x3(_ptr(sv["x1s"]), 2 * H, _ptr(lw["wi"]), 2 * H, _ptr(lw["bi"]), None, 0, _ptr(sv["u"]), 4 * H,
                      M, 4 * H, H, F32, 0, st)
        L.gelu_rows(_ptr(sv["u"]), 0, 4 * H, M, 4 * H, _ptr(sv["inters"]), 8 * H, 4 * H, st)
        if p_h > 0:
            L.gemm_bf16x3(_ptr(sv["inters"]), 8 * H, _ptr(lw["wo2"]), 8 * H, _ptr(lw["bo2"]), None, 0,
                          _ptr(sv["h2"]), H, M, H, 4 * H, F32, 0, st)
            site2 = self._site(sname + ".h2")
            if next_xs is not None:
                L.add_ln_dropout(_ptr(sv["h2"]), 0, H, _ptr(sv["x1"]), 0, H, _ptr(lw["ln2g"]), _ptr(lw["ln2b"]),
                                 LN_EPS_BERT, M, H, None, 0, _ptr(out), H, _ptr(next_xs), 2 * H, 1, 0, 0, 0,
                                 _ptr(sv["h2"]), p_h, seed, site2, st)
            else:
                L.add_ln_dropout(_ptr(sv["h2"]), 0, H, _ptr(sv["x1"]), 0, H, _ptr(lw["ln2g"]), _ptr(lw["ln2b"]),
                                 LN_EPS_BERT, M, H, _ptr(tanh_base), H, _ptr(out), H, _ptr(out16), H, 0, remap[0],
                                 remap[1], remap[2], _ptr(sv["h2"]), p_h, seed, site2, st)
            return
        L.gemm_bf16x3(_ptr(sv["inters"]), 8 * H, _ptr(lw["wo2"]), 8 * H, _ptr(lw["bo2"]), _ptr(sv["x1"]), H,
                      _ptr(sv["h2"]), H, M, H, 4 * H, F32 | RES, 0, st)
        if next_xs is not None:
            L.add_ln_split(_ptr(sv["h2"]), 0, H, None, 0, 0, _ptr(lw["ln2g"]), _ptr(lw["ln2b"]), LN_EPS_BERT, M, H,
                           None, 0, _ptr(out), H, _ptr(next_xs), 2 * H, 0, 0, 0, st)
        else:
            L.add_ln(_ptr(sv["h2"]), 0, H, None, 0, 0, _ptr(lw["ln2g"]), _ptr(lw["ln2b"]), LN_EPS_BERT, M, H,
                     _ptr(tanh_base), H, _ptr(out), H, _ptr(out16), H, remap[0], remap[1], remap[2], st)

    def _bf_layer_fwd(self, L, lw, x, ldx, qkv, sv, M, attn, st, drop=None):
        """One answer-transformer layer over M rows; `attn(qkv, ctx)` enqueues the attention of these rows.
        drop = (p_hidden, seed, site name) of the two hidden dropouts (the attention's is inside `attn`)."""
        p_h, seed, sname = drop if drop is not None else (0.0, 0, "")
        if qkv is None:
            qkv = sv["qkv"]
            L.gemm_bf16(_ptr(x), ldx, _ptr(lw["wqkv"]), H, _ptr(lw["bqkv"]), None, 0, _ptr(qkv), 3 * H, M, 3 * H, H, 0, 0, st)
        attn(qkv, sv["ctx"])
        if p_h > 0:
            L.gemm_bf16(_ptr(sv["ctx"]), H, _ptr(lw["wo"]), H, _ptr(lw["bo"]), None, 0, _ptr(sv["h1"]), H, M, H, H, 0, 0, st)
            L.add_ln_dropout(_ptr(sv["h1"]), 1, H, _ptr(x), 1, ldx, _ptr(lw["ln1g"]), _ptr(lw["ln1b"]), LN_EPS_BERT, M, H,
                             None, 0, None, 0, _ptr(sv["x1"]), H, 0, 0, 0, 0, _ptr(sv["h1"]), p_h, seed,
                             self._site(sname + ".h1"), st)
        else:
            L.gemm_bf16(_ptr(sv["ctx"]), H, _ptr(lw["wo"]), H, _ptr(lw["bo"]), _ptr(x), ldx, _ptr(sv["h1"]), H, M, H, H, 0, 0, st)
            L.add_ln(_ptr(sv["h1"]), 1, H, None, 0, 0, _ptr(lw["ln1g"]), _ptr(lw["ln1b"]), LN_EPS_BERT, M, H, None, 0, None, 0,
                     _ptr(sv["x1"]), H, 0, 0, 0, st)
        L.gemm_bf16(_ptr(sv["x1"]), H, _ptr(lw["wi"]), H, _ptr(lw["bi"]), None, 0, _ptr(sv["u"]), 4 * H, M, 4 * H, H, 0, 0, st)
        L.gelu_rows(_ptr(sv["u"]), 1, 4 * H, M, 4 * H, _ptr(sv["inter"]), 4 * H, 0, st)
        if p_h > 0:
            L.gemm_bf16(_ptr(sv["inter"]), 4 * H, _ptr(lw["wo2"]), 4 * H, _ptr(lw["bo2"]), None, 0, _ptr(sv["h2"]), H,
                        M, H, 4 * H, 0, 0, st)
            L.add_ln_dropout(_ptr(sv["h2"]), 1, H, _ptr(sv["x1"]), 1, H, _ptr(lw["ln2g"]), _ptr(lw["ln2b"]), LN_EPS_BERT,
                             M, H, None, 0, None, 0, _ptr(sv["out"]), H, 0, 0, 0, 0, _ptr(sv["h2"]), p_h, seed,
                             self._site(sname + ".h2"), st)
        else:
            L.gemm_bf16(_ptr(sv["inter"]), 4 * H, _ptr(lw["wo2"]), 4 * H, _ptr(lw["bo2"]), _ptr(sv["x1"]), H, _ptr(sv["h2"]), H,
                        M, H, 4 * H, 0, 0, st)
            L.add_ln(_ptr(sv["h2"]), 1, H, None, 0, 0, _ptr(lw["ln2g"]), _ptr(lw["ln2b"]), LN_EPS_BERT, M, H, None, 0, None, 0,
                     _ptr(sv["out"]), H, 0, 0, 0, st)
        return qkv

    # ------------------------------------------------------------------ forward
    def forward(self, inp):
        m = self.model
        L = _lib.get_lib()
        dev = self.dev
        B, Lt = inp["text"].shape
        if self.is_t2s:
            F, O = inp["video_feat"].shape[1], inp["ocr_mask"].shape[1]
        else:
            F, O = m._sv_dims(inp)            # rows of the object part (1 for M4C), OCR slots
        T = inp["train_prev_inds"].shape[1]
        V = m.classifier.module.weight.shape[0]
        Le = Lt + F + O
        P = m._pack(dev)
        variants = self.variants
        mws = m._workspace(B, dev, dict(Lt=Lt, F=F, O=O, T=T, V=V, k_obj_pad=P["k_obj_pad"], k_ocr_pad=P["k_ocr_pad"],
                                        variants=variants))
        ws = self._train_ws(B, Lt, F, O, T, V, P["k_obj_pad"], P["k_ocr_pad"])
        st = torch.cuda.current_stream(dev).cuda_stream
        f = P["f32"]
        Me, Md, Mt = B * Le, B * T, B * Lt
        seed = self._step_seed()
        pt_h, pt_a = self._p("text", 0), self._p("text", 1)
        pq_h, pq_a = (self._p("qtv", 0), self._p("qtv", 1)) if self.is_t2s else (0.0, 0.0)
        pm_h, pm_a = self._p("mmt", 0), self._p("mmt", 1)
        p_obj, p_ocr = self._p("obj"), self._p("ocr")

        # ---- masks and key lists (as the eval path)
        if self.is_t2s:
            L.mask_prep(_ptr(inp["text_len"]), _ptr(inp["frame_mask"]), _ptr(inp["ocr_mask"]), B, Lt, F, O, _ptr(mws["jm_ref"]), st)
            L.mask_prep(_ptr(inp["text_len"]), None, None, B, Lt, 0, 0, _ptr(mws["jm_txt"]), st)
            L.build_keys(_ptr(mws["jm_txt"]), B, Lt, _ptr(mws["keys_txt"]), _ptr(mws["nk_txt"]), Lt, st)
            L.build_keys(_ptr(mws["jm_ref"]), B, Le, _ptr(mws["keys"]["ref"]), _ptr(mws["nk"]["ref"]), Le, st)
        else:
            m._sv_masks(L, mws, inp, B, Lt, F, O, st)

        # ---- TextBert
        e = "text_bert.embeddings."
        L.bert_embed_ln(_ptr(inp["text"]), Mt, Lt, H, _ptr(f[e + "word_embeddings.weight"]),
                        _ptr(f[e + "position_embeddings.weight"]), _ptr(f[e + "token_type_embeddings.weight"]),
                        _ptr(f[e + "LayerNorm.weight"]), _ptr(f[e + "LayerNorm.bias"]), LN_EPS_BERT, _ptr(ws["xt"]), H, st)
        x = ws["xt"]
        n = len(P["text"])
        if pt_h > 0:       # BertEmbeddings: LayerNorm -> dropout
            L.dropout_rows(_ptr(x), 0, H, Mt, H, 0, 0, 0, pt_h, seed, self._site("text.emb"), st)
        L.split_bf16(_ptr(x), H, Mt, H, H, _ptr(ws["text"][0]["xs"]), 2 * H, 0, 0, 0, st)
        for i, lw in enumerate(P["text"]):
            sv, last = ws["text"][i], i == n - 1
            self._x3_layer_fwd(L, lw, x, sv, Mt, Lt, mws["keys_txt"], mws["nk_txt"], Lt, st,
                               out=mws["J0"] if last else sv["out"], remap=(Lt, Le, 0) if last else (0, 0, 0),
                               next_xs=None if last else ws["text"][i + 1]["xs"],
                               drop=(pt_h, pt_a, seed, "text.%d" % i))
            x = sv["out"]
        # ---- obj / OCR encoders (model code; h_obj / h_ocr / a_*_s stay in the model workspace until backward)
        if self.is_t2s:
            enc_inp, has_ids = inp, True
            m._encode_obj_ocr(L, P, mws, inp, B, Lt, F, O, Le, st)
        else:
            enc_inp, m4c = m._sv_encoder_inputs(inp)
            has_ids = not m4c
            m._encode_obj_ocr(L, P, mws, enc_inp, B, Lt, F, O, Le, st, m4c=m4c)
        # obj_drop / ocr_drop (t2s.py:214,253; m4c.py:206,247) on the encoded rows of the joint buffer
        if p_obj > 0:
            L.dropout_rows(_ptr(mws["J0"]), 0, H, B * F, H, F, Le, Lt, p_obj, seed, self._site("obj"), st)
        if p_ocr > 0:
            L.dropout_rows(_ptr(mws["J0"]), 0, H, B * O, H, O, Le, Lt + F, p_ocr, seed, self._site("ocr"), st)
        if self.is_t2s:
            # ---- QTV
            x, n = mws["J0"], len(P["qtv"])
            L.split_bf16(_ptr(x), H, Me, H, H, _ptr(ws["qtv"][0]["xs"]), 2 * H, 0, 0, 0, st)
            for i, lw in enumerate(P["qtv"]):
                sv, last = ws["qtv"][i], i == n - 1
                self._x3_layer_fwd(L, lw, x, sv, Me, Le, mws["keys"]["ref"], mws["nk"]["ref"], Le, st,
                                   out=mws["J1"] if last else sv["out"], tanh_base=mws["J0"] if last else None,
                                   out16=mws["X16"] if last else None, next_xs=None if last else ws["qtv"][i + 1]["xs"],
                                   drop=(pq_h, pq_a, seed, "qtv.%d" % i))
                x = sv["out"]
            # ---- grounding (no gradient: emits constant masks, SURVEY hard part 9)
            ground_frame, ground_box, _, _, _ = m._grounding(L, P, mws, inp, B, Lt, F, O, O // F, Le, dev, st)
            jm = {"ref": mws["jm_ref"], "pos": mws["jm_pos"], "neg": mws["jm_neg"]}
        else:
            # no QTV: the encoders' output feeds the answer transformer (m4c.py:255-273, t5vitevqa.py, gt_box.py:298-299);
            # the model's grounding hook (post-hoc attention, no gradient) fixes the variant's key list
            mws["J1"].copy_(mws["J0"])
            L.cast_rows_bf16(_ptr(mws["J0"]), H, Me, H, _ptr(mws["X16"]), H, 0, 0, 0, st)
            ground_frame, ground_box = m._sv_ground(L, P, mws, inp, B, Lt, F, O, Le, dev, st)
            jm = {variants[0]: mws["jm_" + variants[0]]}

        # ---- answer transformer, teacher forced (reference t2s.py:288-314)
        N = V + O
        scores = {v: torch.empty(B, T, N, device=dev, dtype=torch.float32) for v in variants}
        layers = P["mmt"]
        prev = inp["train_prev_inds"].contiguous()
        pp = "mmt.prev_pred_embeddings."
        ocr_row0 = Lt + F
        L.gemm_bf16(_ptr(mws["X16"]), H, _ptr(layers[0]["wqkv"]), H, _ptr(layers[0]["bqkv"]), None, 0, _ptr(mws["qkv0"]),
                    3 * H, Me, 3 * H, H, 0, 0, st)
        enc_qkv = {}
        for v in variants:
            keys, nk = mws["keys"][v], mws["nk"][v]

            def enc_attn(qkv, ctx, li, keys=keys, nk=nk, v=v):
                # one dropout site and one statistics buffer per (variant, layer): encoder and decoder rows are one
                # virtual sequence of Le + T query rows in the backward
                L.attn_tc_dropout(_ptr(qkv), 3 * H, 0, B, Le, H, 12, _ptr(keys), _ptr(nk), Le, _ptr(ctx), H, pm_a,
                                  seed, self._site("mmt.%s.%d.attn" % (v, li)), _ptr(ws["enc"][v][li]["stats"]), Le + T, st)

            x = mws["X16"]
            enc_qkv[v] = []
            for li, lw in enumerate(layers):
                sv = ws["enc"][v][li]
                q = self._bf_layer_fwd(L, lw, x, H, mws["qkv0"] if li == 0 else None, sv, Me,
                                       lambda qkv, ctx, li=li: enc_attn(qkv, ctx, li), st,
                                       drop=(pm_h, seed, "mmt.%s.%d.enc" % (v, li)))
                enc_qkv[v].append(q)
                x = sv["out"]
            L.gemm_bf16(_ptr(x), H, _ptr(P["w_ptr_k"]), H, _ptr(f["ocr_ptr_net.key.bias"]), None, 0, _ptr(ws["keyp"][v]), H,
                        Me, H, H, 0, 0, st)
            # decoder rows
            L.prev_embed(_ptr(prev), T, B, 0, T, T, V, H, _ptr(f["classifier.module.weight"]),
                         mws["J1"].data_ptr() + ocr_row0 * H * 4, Le * H, H,
                         _ptr(f[pp + "position_embeddings.weight"]), _ptr(f[pp + "token_type_embeddings.weight"]),
                         _ptr(f[pp + "ans_layer_norm.weight"]), _ptr(f[pp + "ans_layer_norm.bias"]),
                         _ptr(f[pp + "ocr_layer_norm.weight"]), _ptr(f[pp + "ocr_layer_norm.bias"]),
                         _ptr(f[pp + "emb_layer_norm.weight"]), _ptr(f[pp + "emb_layer_norm.bias"]), LN_EPS_BERT,
                         _ptr(ws["xd"][v]), None, H, O, st)
            if pm_h > 0:       # PrevPredEmbeddings.emb_dropout (t2s.py:720)
                L.dropout_rows(_ptr(ws["xd"][v]), 1, H, Md, H, 0, 0, 0, pm_h, seed, self._site("mmt.%s.prev" % v), st)
            x = ws["xd"][v]
            for li, lw in enumerate(layers):
                sv = ws["dec"][v][li]
                qe = enc_qkv[v][li]

                def dec_attn(qkv, ctx, qe=qe, keys=keys, nk=nk, li=li, v=v):
                    L.attn_dec_dropout(_ptr(qe), 3 * H, Le, _ptr(qkv), 3 * H, T, B, H, 12, _ptr(keys), _ptr(nk), Le,
                                       0, T, _ptr(ctx), H, pm_a, seed, self._site("mmt.%s.%d.attn" % (v, li)),
                                       _ptr(ws["enc"][v][li]["stats"]), st)

                self._bf_layer_fwd(L, lw, x, H, None, sv, Md, dec_attn, st, drop=(pm_h, seed, "mmt.%s.%d.dec" % (v, li)))
                x = sv["out"]
            sc = scores[v]
            L.gemm_bf16(_ptr(x), H, _ptr(P["w_cls"]), H, _ptr(f["classifier.module.bias"]), None, 0, _ptr(sc), N,
                        Md, V, H, _lib.GEMM_OUT_F32, 0, st)
            L.gemm_bf16(_ptr(x), H, _ptr(P["w_ptr_q"]), H, _ptr(f["ocr_ptr_net.query.bias"]), None, 0, _ptr(ws["qd"][v]), H,
                        Md, H, H, 0, 0, st)
            L.ptr_score(_ptr(ws["qd"][v]), H, B, T, 0, T, ws["keyp"][v].data_ptr() + ocr_row0 * H * 2, Le * H, H, O, H,
                        jm[v].data_ptr() + ocr_row0 * 4, Le, _ptr(sc), N, V, st)
        self.saved = dict(inp=inp, enc_inp=enc_inp, has_ids=has_ids, dims=(B, Lt, F, O, T, V, Le), enc_qkv=enc_qkv, prev=prev,
                          seed=seed, p=dict(text=(pt_h, pt_a), qtv=(pq_h, pq_a), mmt=(pm_h, pm_a), obj=p_obj, ocr=p_ocr))
        self.fwd_gen += 1
        self.ground = (ground_frame, ground_box)
        return [scores[v] for v in self.out_variants]

    # ------------------------------------------------------------------ backward building blocks
    def _wgrad(self, L, G, ldg, X, ldx, dW_ptr, ldd, rows, Pn, Qn, st):
        L.gemm_wgrad_bf16(_ptr(G) if torch.is_tensor(G) else G, ldg, _ptr(X) if torch.is_tensor(X) else X, ldx, dW_ptr,
                          ldd, rows, Pn, Qn, 0, st)

    def _layer_bwd_pre_attn(self, L, wt, lw, pre, sv, M, dy, sc, st, x3, drop=None):
        """LN2 -> FFN -> LN1 -> attention-out of one layer over M rows.  dy: gradient of the layer output (bf16, or the
        (tensor, flags) of the QTV tail); leaves d(context) in sc["dctx"] and the LN1 input gradient in sc["dh1"].
        `x3`: the layer is a grounding-chain layer (fp32 saved pre-activations, bf16 hi|lo operand buffers).
        drop = (p_hidden, seed, site name) of the forward's hidden dropouts: the Linear outputs (dgrad, wgrad, bias) get
        the masked gradient sc["dh*m"], the residual branches the unmasked sc["dh*"]."""
        g = self._g
        hb = 0 if x3 else 1                     # saved pre-LayerNorm sums: fp32 in the grounding chain, bf16 in MMT
        inter, ld_inter = (sv["inters"], 8 * H) if x3 else (sv["inter"], 4 * H)
        x1, ld_x1 = (sv["x1s"], 2 * H) if x3 else (sv["x1"], H)
        ctx, ld_ctx = (sv["ctxs"], 2 * H) if x3 else (sv["ctx"], H)
        dy_t, dy_bf16, dy_map, tanh_out = dy
        p_h, seed, sname = drop if drop is not None else (0.0, 0, "")
        if p_h > 0:
            g2, g1 = sc["dh2m"], sc["dh1m"]
            L.ln_bwd_dropout(_ptr(sv["h2"]), hb, H, _ptr(dy_t), dy_bf16, H, dy_map[0], dy_map[1], dy_map[2], _ptr(lw["ln2g"]),
                             _ptr(lw["ln2b"]), LN_EPS_BERT, M, H, tanh_out, _ptr(sc["dh2"]), 1, H,
                             g(pre + "output.LayerNorm.weight"), g(pre + "output.LayerNorm.bias"),
                             g(pre + "output.dense.bias"), _ptr(g2), p_h, seed, self._site(sname + ".h2"), st)
        else:
            g2, g1 = sc["dh2"], sc["dh1"]
            L.ln_bwd(_ptr(sv["h2"]), hb, H, _ptr(dy_t), dy_bf16, H, dy_map[0], dy_map[1], dy_map[2], _ptr(lw["ln2g"]),
                     _ptr(lw["ln2b"]), LN_EPS_BERT, M, H, tanh_out, _ptr(sc["dh2"]), 1, H,
                     g(pre + "output.LayerNorm.weight"), g(pre + "output.LayerNorm.bias"), g(pre + "output.dense.bias"), st)
        L.gemm_bf16(_ptr(g2), H, _ptr(wt["wo2T"]), H, None, _ptr(sv["u"]), 4 * H, _ptr(sc["du"]), 4 * H, M, 4 * H, H,
                    _lib.GEMM_DGELU | (_lib.GEMM_RES_F32 if x3 else 0), 0, st)
        self._wgrad(L, g2, H, inter, ld_inter, g(pre + "output.dense.weight"), 4 * H, M, H, 4 * H, st)
        L.colsum(_ptr(sc["du"]), 1, 4 * H, M, 4 * H, g(pre + "intermediate.dense.bias"), st)
        L.gemm_bf16(_ptr(sc["du"]), 4 * H, _ptr(wt["wiT"]), 4 * H, None, _ptr(sc["dh2"]), H, _ptr(sc["dx1"]), H, M, H, 4 * H,
                    0, 0, st)
        self._wgrad(L, sc["du"], 4 * H, x1, ld_x1, g(pre + "intermediate.dense.weight"), H, M, 4 * H, H, st)
        if p_h > 0:
            L.ln_bwd_dropout(_ptr(sv["h1"]), hb, H, _ptr(sc["dx1"]), 1, H, 0, 0, 0, _ptr(lw["ln1g"]), _ptr(lw["ln1b"]),
                             LN_EPS_BERT, M, H, 0, _ptr(sc["dh1"]), 1, H, g(pre + "attention.output.LayerNorm.weight"),
                             g(pre + "attention.output.LayerNorm.bias"), g(pre + "attention.output.dense.bias"),
                             _ptr(g1), p_h, seed, self._site(sname + ".h1"), st)
        else:
            L.ln_bwd(_ptr(sv["h1"]), hb, H, _ptr(sc["dx1"]), 1, H, 0, 0, 0, _ptr(lw["ln1g"]), _ptr(lw["ln1b"]), LN_EPS_BERT,
                     M, H, 0, _ptr(sc["dh1"]), 1, H, g(pre + "attention.output.LayerNorm.weight"),
                     g(pre + "attention.output.LayerNorm.bias"), g(pre + "attention.output.dense.bias"), st)
        L.gemm_bf16(_ptr(g1), H, _ptr(wt["woT"]), H, None, None, 0, _ptr(sc["dctx"]), H, M, H, H, 0, 0, st)
        self._wgrad(L, g1, H, ctx, ld_ctx, g(pre + "attention.output.dense.weight"), H, M, H, H, st)

    def _attn_bwd(self, L, args, p_a, seed, sname, st, stats):
        """Attention backward of one layer: the forward's attention-probability dropout recomputed from (seed, site)
        (p_a may be 0), the rows' log2-sum-exp taken from `stats`, where the forward kernels left it."""
        L.attn_bwd_dropout(*args, p_a, seed, self._site(sname + ".attn"), _ptr(stats), st)

    def _layer_bwd_post_attn(self, L, wt, pre, x, ldx, M, sc, dx_out, st):
        """q|k|v projection backward: sc["dqkv"] -> dx_out (+ the residual gradient sc["dh1"]); weight / bias grads."""
        g = self._g
        L.colsum(_ptr(sc["dqkv"]), 1, 3 * H, M, 3 * H, g(pre + "attention.self.query.bias"), st)
        L.gemm_bf16(_ptr(sc["dqkv"]), 3 * H, _ptr(wt["wqkvT"]), 3 * H, None, _ptr(sc["dh1"]), H, _ptr(dx_out), H, M, H, 3 * H,
                    0, 0, st)
        self._wgrad(L, sc["dqkv"], 3 * H, x, ldx, g(pre + "attention.self.query.weight"), H, M, 3 * H, H, st)

    # ------------------------------------------------------------------ backward
    def backward(self, dscores):
        m = self.model
        L = _lib.get_lib()
        dev = self.dev
        sv_all = self.saved
        if sv_all is None:
            raise RuntimeError("backward called without a training forward")
        inp = sv_all["inp"]
        enc_inp, has_ids = sv_all["enc_inp"], sv_all["has_ids"]
        B, Lt, F, O, T, V, Le = sv_all["dims"]
        P = m._packed
        W = self._wt_pack()
        variants = self.variants
        mws = m._workspace(B, dev, dict(Lt=Lt, F=F, O=O, T=T, V=V, k_obj_pad=P["k_obj_pad"], k_ocr_pad=P["k_ocr_pad"],
                                        variants=variants))
        ws = self._train_ws(B, Lt, F, O, T, V, P["k_obj_pad"], P["k_ocr_pad"])
        st = torch.cuda.current_stream(dev).cuda_stream
        f = P["f32"]
        g = self._g
        Me, Md, Mt = B * Le, B * T, B * Lt
        N, Np = V + O, ws["Np"]
        ocr_row0 = Lt + F
        layers = P["mmt"]
        n_mmt = len(layers)
        self.flat_grad[:self.live_end].zero_()
        self._begin_overlapped_reduce()
        live_v = [v for v in variants if dscores.get(v) is not None]
        enc_sc = dict(dh2=ws["dh2"], du=ws["du"], dx1=ws["dx1"], dh1=ws["dh1"], dctx=ws["dctx"], dqkv=ws["dqkv"],
                      dh2m=ws["dh2m"], dh1m=ws["dh1m"])
        dec_sc = dict(dh2=ws["dh2_d"], du=ws["du_d"], dx1=ws["dx1_d"], dh1=ws["dh1_d"], dctx=ws["dctx_d"], dqkv=ws["dqkv_d"],
                      dh2m=ws["dh2m_d"], dh1m=ws["dh1m_d"])
        seed = sv_all["seed"]
        pt_h, pt_a = sv_all["p"]["text"]
        pq_h, pq_a = sv_all["p"]["qtv"]
        pm_h, pm_a = sv_all["p"]["mmt"]
        p_obj, p_ocr = sv_all["p"]["obj"], sv_all["p"]["ocr"]
        pp = "mmt.prev_pred_embeddings."
        first_variant = True
        for v in variants:
            dS = dscores.get(v)
            if dS is None:
                continue
            dS = dS.contiguous()
            keys, nk = mws["keys"][v], mws["nk"][v]
            xdec_out = ws["dec"][v][n_mmt - 1]["out"]
            xenc_out = ws["enc"][v][n_mmt - 1]["out"]
            # ---- heads
            L.cast_rows_bf16(_ptr(dS), N, Md, N, _ptr(ws["dS16"]), Np, 0, 0, 0, st)
            L.colsum(_ptr(dS), 0, N, Md, V, g("classifier.module.bias"), st)
            L.gemm_bf16(_ptr(ws["dS16"]), Np, _ptr(W["clsT"]), W["clsT"].shape[1], None, None, 0, _ptr(ws["dy_d"]), H,
                        Md, H, V, 0, 0, st)
            self._wgrad(L, ws["dS16"], Np, xdec_out, H, g("classifier.module.weight"), H, Md, V, H, st)
            L.ptr_score_bwd(_ptr(dS), N, B, T, V, _ptr(ws["qd"][v]), H, ws["keyp"][v].data_ptr() + ocr_row0 * H * 2, Le * H, H,
                            O, H, _ptr(ws["dq"]), H, ws["dkeyp"].data_ptr() + ocr_row0 * H * 2, Le * H, H, st)
            L.colsum(_ptr(ws["dq"]), 1, H, Md, H, g("ocr_ptr_net.query.bias"), st)
            L.gemm_bf16(_ptr(ws["dq"]), H, _ptr(W["ptr_qT"]), H, None, _ptr(ws["dy_d"]), H, _ptr(ws["dy_d"]), H, Md, H, H,
                        0, 0, st)
            self._wgrad(L, ws["dq"], H, xdec_out, H, g("ocr_ptr_net.query.weight"), H, Md, H, H, st)
            L.colsum(_ptr(ws["dkeyp"]), 1, H, Me, H, g("ocr_ptr_net.key.bias"), st)
            L.gemm_bf16(_ptr(ws["dkeyp"]), H, _ptr(W["ptr_kT"]), H, None, None, 0, _ptr(ws["dy"]), H, Me, H, H, 0, 0, st)
            self._wgrad(L, ws["dkeyp"], H, xenc_out, H, g("ocr_ptr_net.key.weight"), H, Me, H, H, st)
            # ---- layers, last to first
            dy_e, dy_d = ws["dy"], ws["dy_d"]
            alt_e, alt_d = ws["dy2"], ws["dy2_d"]
            for li in range(n_mmt - 1, -1, -1):
                lw, wt = layers[li], W["mmt"][li]
                pre = "mmt.encoder.layer.%d." % li
                sve, svd = ws["enc"][v][li], ws["dec"][v][li]
                qkv_e = sv_all["enc_qkv"][v][li]
                x_e = mws["X16"] if li == 0 else ws["enc"][v][li - 1]["out"]
                x_d = ws["xd"][v] if li == 0 else ws["dec"][v][li - 1]["out"]
                self._layer_bwd_pre_attn(L, wt, lw, pre, svd, Md, (dy_d, 1, (0, 0, 0), 0), dec_sc, st, x3=False,
                                         drop=(pm_h, seed, "mmt.%s.%d.dec" % (v, li)))
                self._layer_bwd_pre_attn(L, wt, lw, pre, sve, Me, (dy_e, 1, (0, 0, 0), 0), enc_sc, st, x3=False,
                                         drop=(pm_h, seed, "mmt.%s.%d.enc" % (v, li)))
                self._attn_bwd(L, (_ptr(qkv_e), 3 * H, _ptr(svd["qkv"]), 3 * H, _ptr(sve["ctx"]), H, _ptr(svd["ctx"]), H,
                                   _ptr(enc_sc["dctx"]), H, _ptr(dec_sc["dctx"]), H, _ptr(enc_sc["dqkv"]), 3 * H,
                                   _ptr(dec_sc["dqkv"]), 3 * H, B, Le, T, H, 12, _ptr(keys), _ptr(nk), Le, Le,
                                   _ptr(ws["attn_ws"])), pm_a, seed, "mmt.%s.%d" % (v, li), st, sve["stats"])
                self._layer_bwd_post_attn(L, wt, pre, x_d, H, Md, dec_sc, alt_d, st)
                self._layer_bwd_post_attn(L, wt, pre, x_e, H, Me, enc_sc, alt_e, st)
                dy_e, alt_e = alt_e, dy_e
                dy_d, alt_d = alt_d, dy_d
                if v == live_v[-1]:        # the last variant's pass over this layer: its weight gradients are final
                    self._bucket_ready(pre)
            # ---- encoder-input gradient of this variant into dJ1; decoder input through PrevPredEmbeddings
            L.rows_add(_ptr(dy_e), None, None, H, Me, H, _ptr(ws["dJ"]), H, 0, 0, 0, 0 if first_variant else 1, st)
            first_variant = False
            if pm_h > 0:       # emb_dropout of the forward: the same mask on the gradient of the decoder input rows
                L.dropout_rows(_ptr(dy_d), 1, H, Md, H, 0, 0, 0, pm_h, seed, self._site("mmt.%s.prev" % v), st)
            L.prev_embed_bwd(_ptr(dy_d), H, _ptr(sv_all["prev"]), T, B, T, V, H, _ptr(f["classifier.module.weight"]),
                             mws["J1"].data_ptr() + ocr_row0 * H * 4, Le * H, H,
                             _ptr(f[pp + "position_embeddings.weight"]), _ptr(f[pp + "token_type_embeddings.weight"]),
                             _ptr(f[pp + "ans_layer_norm.weight"]), _ptr(f[pp + "ocr_layer_norm.weight"]),
                             _ptr(f[pp + "emb_layer_norm.weight"]), LN_EPS_BERT, g("classifier.module.weight"),
                             ws["dJ"].data_ptr() + ocr_row0 * H * 4, g(pp + "position_embeddings.weight"),
                             g(pp + "token_type_embeddings.weight"), g(pp + "ans_layer_norm.weight"),
                             g(pp + "ans_layer_norm.bias"), g(pp + "ocr_layer_norm.weight"), g(pp + "ocr_layer_norm.bias"),
                             g(pp + "emb_layer_norm.weight"), g(pp + "emb_layer_norm.bias"), O, st)
        if first_variant:
            raise RuntimeError("no score gradient reached the model")
        self._bucket_ready("mmt.prev_pred_embeddings.")
        self._bucket_ready("classifier.")
        self._bucket_ready("ocr_ptr_net.")

        # ---- QTV: J1 = J0 + tanh(LN2_last(.)): dJ holds dJ1; after the loop dJ += dx(layer 0) = dJ0
        # (single-variant models have no QTV: dJ already is dJ0)
        qtv = P["qtv"] if self.is_t2s else []
        dy = (ws["dJ"], 0, (0, 0, 0), 1)
        for li in range(len(qtv) - 1, -1, -1):
            lw, wt, sv = qtv[li], W["qtv"][li], ws["qtv"][li]
            pre = "TransLayer.encoder.layer.%d." % li
            self._layer_bwd_pre_attn(L, wt, lw, pre, sv, Me, dy, enc_sc, st, x3=True, drop=(pq_h, seed, "qtv.%d" % li))
            self._attn_bwd(L, (_ptr(sv["qkvs"]), 6 * H, None, 0, _ptr(sv["ctxs"]), 2 * H, None, 0, _ptr(enc_sc["dctx"]), H,
                               None, 0, _ptr(enc_sc["dqkv"]), 3 * H, None, 0, B, Le, 0, H, 12, _ptr(mws["keys"]["ref"]),
                               _ptr(mws["nk"]["ref"]), Le, Le, _ptr(ws["attn_ws"])), pq_a, seed, "qtv.%d" % li, st, sv["stats"])
            self._layer_bwd_post_attn(L, wt, pre, sv["xs"], 2 * H, Me, enc_sc, ws["dy"], st)
            dy = (ws["dy"], 1, (0, 0, 0), 0)
            self._bucket_ready(pre)
        if qtv:
            L.rows_add(_ptr(ws["dy"]), None, None, H, Me, H, _ptr(ws["dJ"]), H, 0, 0, 0, 1, st)      # dJ = dJ0

        # ---- obj_drop / ocr_drop of the forward: the same masks on the gradient rows of the joint buffer
        if p_obj > 0:
            L.dropout_rows(_ptr(ws["dJ"]), 0, H, B * F, H, F, Le, Lt, p_obj, seed, self._site("obj"), st)
        if p_ocr > 0:
            L.dropout_rows(_ptr(ws["dJ"]), 0, H, B * O, H, O, Le, Lt + F, p_ocr, seed, self._site("ocr"), st)
        # ---- obj encoder: J0[obj rows] = LN(W a + b)
        L.ln_bwd(_ptr(mws["h_obj"]), 0, H, _ptr(ws["dJ"]), 0, H, F, Le, Lt, _ptr(f["obj_feat_layer_norm.weight"]),
                 _ptr(f["obj_feat_layer_norm.bias"]), LN_EPS_EMBED, B * F, H, 0, _ptr(ws["dh_obj"]), 1, H,
                 g("obj_feat_layer_norm.weight"), g("obj_feat_layer_norm.bias"), g("linear_obj_feat_to_mmt_in.bias"), st)
        kpo, kpc = P["k_obj_pad"], P["k_ocr_pad"]
        ko, kc = P["k_obj"], P["k_ocr"]
        ws["dw_obj"].zero_()
        self._wgrad(L, ws["dh_obj"], H, mws["a_obj_s"], 2 * kpo, _ptr(ws["dw_obj"]), kpo, B * F, H, kpo, st)
        self.grad("linear_obj_feat_to_mmt_in.weight").add_(ws["dw_obj"][:, :ko])
        if has_ids:         # gradient of the frame-id embedding (M4C's object token has none)
            L.gemm_bf16(_ptr(ws["dh_obj"]), H, W["objT"].data_ptr() + (ko - 50) * H * 2, H, None, None, 0, _ptr(ws["d_id_obj"]), 52,
                        B * F, 50, H, _lib.GEMM_OUT_F32, 0, st)
            L.embed_scatter_add(_ptr(ws["d_id_obj"]), 0, 52, 0, 50, _ptr(enc_inp["frame_id"]), B * F, -1,
                                g("frame_embeddings.weight"), 50, st)
        # ---- OCR encoder: J0[ocr rows] = LN(W a + b) + LN(W2 bbox + b2)
        L.ocr_finish_bwd(_ptr(mws["h_ocr"]), H, _ptr(enc_inp["ocr_bbox_coordinates"]), _ptr(f["linear_ocr_bbox_to_mmt_in.weight"]),
                         _ptr(f["linear_ocr_bbox_to_mmt_in.bias"]), _ptr(f["ocr_feat_layer_norm.weight"]),
                         _ptr(f["ocr_bbox_layer_norm.weight"]), LN_EPS_EMBED, B * O, H, _ptr(ws["dJ"]), H, O, Le, Lt + F,
                         _ptr(ws["dh_ocr"]), H, _ptr(ws["dc_ocr"]), H, g("ocr_feat_layer_norm.weight"),
                         g("ocr_feat_layer_norm.bias"), g("ocr_bbox_layer_norm.weight"), g("ocr_bbox_layer_norm.bias"),
                         g("linear_ocr_feat_to_mmt_in.bias"), g("linear_ocr_bbox_to_mmt_in.weight"),
                         g("linear_ocr_bbox_to_mmt_in.bias"), st)
        ws["dw_ocr"].zero_()
        self._wgrad(L, ws["dh_ocr"], H, mws["a_ocr_s"], 2 * kpc, _ptr(ws["dw_ocr"]), kpc, B * O, H, kpc, st)
        self.grad("linear_ocr_feat_to_mmt_in.weight").add_(ws["dw_ocr"][:, :kc])
        if has_ids:
            L.gemm_bf16(_ptr(ws["dh_ocr"]), H, W["ocrT"].data_ptr() + (kc - 100) * H * 2, H, None, None, 0, _ptr(ws["d_id_ocr"]), 100,
                        B * O, 100, H, _lib.GEMM_OUT_F32, 0, st)
            L.embed_scatter_add(_ptr(ws["d_id_ocr"]), 0, 100, 0, 50, _ptr(enc_inp["temporal_id"]), B * O, -1,
                                g("temporal_position_embeddings.weight"), 50, st)
            L.embed_scatter_add(_ptr(ws["d_id_ocr"]), 0, 100, 50, 50, _ptr(enc_inp["track_id"]), B * O, -1,
                                g("track_position_embeddings.weight"), 50, st)
        # ---- TextBert
        text = P["text"]
        dy = (ws["dJ"], 0, (Lt, Le, 0), 0)
        txt_sc = {k: t[:Mt] for k, t in enc_sc.items()}
        dx_t = ws["dy2"][:Mt]
        for li in range(len(text) - 1, -1, -1):
            lw, wt, sv = text[li], W["text"][li], ws["text"][li]
            pre = "text_bert.encoder.layer.%d." % li
            self._layer_bwd_pre_attn(L, wt, lw, pre, sv, Mt, dy, txt_sc, st, x3=True, drop=(pt_h, seed, "text.%d" % li))
            self._attn_bwd(L, (_ptr(sv["qkvs"]), 6 * H, None, 0, _ptr(sv["ctxs"]), 2 * H, None, 0, _ptr(txt_sc["dctx"]), H,
                               None, 0, _ptr(txt_sc["dqkv"]), 3 * H, None, 0, B, Lt, 0, H, 12, _ptr(mws["keys_txt"]),
                               _ptr(mws["nk_txt"]), Lt, Lt, _ptr(ws["attn_ws"])), pt_a, seed, "text.%d" % li, st, sv["stats"])
            self._layer_bwd_post_attn(L, wt, pre, sv["xs"], 2 * H, Mt, txt_sc, dx_t, st)
            dy = (dx_t, 1, (0, 0, 0), 0)
            self._bucket_ready(pre)
        if pt_h > 0:           # BertEmbeddings' dropout
            L.dropout_rows(_ptr(dx_t), 1, H, Mt, H, 0, 0, 0, pt_h, seed, self._site("text.emb"), st)
        e = "text_bert.embeddings."
        L.bert_embed_bwd(_ptr(dx_t), H, _ptr(inp["text"]), Mt, Lt, H, _ptr(f[e + "word_embeddings.weight"]),
                         _ptr(f[e + "position_embeddings.weight"]), _ptr(f[e + "token_type_embeddings.weight"]),
                         _ptr(f[e + "LayerNorm.weight"]), LN_EPS_BERT, g(e + "word_embeddings.weight"),
                         g(e + "position_embeddings.weight"), g(e + "token_type_embeddings.weight"),
                         g(e + "LayerNorm.weight"), g(e + "LayerNorm.bias"), st)
        self.saved = None
        if self._reduced_upto is not None:
            self._finish_overlapped_reduce()  # remaining ranges (encoders' linears / id tables, embeddings) + join
        elif self._dist_world() > 1 and self.reduce_in_backward:
            self.all_reduce()                 # one flat NCCL call, still inside loss.backward(): p.grad gets the mean
        return [self.grad(n) for n in self.live_names]

    # ------------------------------------------------------------------ bucketed gradient all-reduce
    def _range_of(self, prefixes):
        """[lo, hi) of the flat buffer covered by the live parameters whose names start with one of `prefixes`;
        they must be contiguous there (they are: `_flatten` sorts by name inside each optimizer group)."""
        key = tuple(prefixes)
        cache = self.__dict__.setdefault("_range_cache", {})
        if key in cache:
            return cache[key]
        names = [n for n in self.live_names if n.startswith(key)]
        if not names:
            cache[key] = None
            return None
        lo = min(self.offsets[n] for n in names)
        hi = max(self.offsets[n] + _ru(self.named[n].numel(), 4) for n in names)
        inside = [n for n in self.live_names if lo <= self.offsets[n] < hi]
        if sorted(inside) != sorted(names):
            raise AssertionError("gradient bucket %s is not contiguous in the flat buffer" % (key,))
        cache[key] = (lo, hi)
        return cache[key]

    def _dist_world(self):
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _bucket_ready(self, *prefixes):
        """The gradients of the parameters under `prefixes` are final on the compute stream: all-reduce that range on
        the communication stream, overlapping whatever the backward still has to do."""
        if self._reduced_upto is None:            # not a distributed, overlapped backward
            return
        r = self._range_of(prefixes)
        if r is None:
            return
        self._ready_ranges.append(r)
        if sum(hi - lo for lo, hi in self._ready_ranges) * 4 >= self.bucket_bytes:
            self._flush_ready()

    def _flush_ready(self):
        """All-reduce the collected final ranges (adjacent ones merged) on the communication stream, behind an event of
        the compute stream."""
        if not self._ready_ranges:
            return
        import torch.distributed as dist
        merged = []
        for lo, hi in sorted(self._ready_ranges):
            if merged and lo <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], hi)
            else:
                merged.append([lo, hi])
        self._ready_ranges = []
        main = torch.cuda.current_stream(self.dev)
        ready = torch.cuda.Event()
        ready.record(main)
        comm = self._comm_stream
        comm.wait_event(ready)
        with torch.cuda.stream(comm):
            for lo, hi in merged:
                if self.time_comm:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(comm)
                dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM)
                if self.time_comm:
                    e1.record(comm)
                    self.comm_events.append(((hi - lo) * 4, e0, e1))
                self._buckets_pending.append((lo, hi))

    def _begin_overlapped_reduce(self):
        self._buckets_pending = []
        self._ready_ranges = []
        self.comm_events = []
        self._reduced_upto = None
        if not self.overlap_allreduce or not self.reduce_in_backward or self._dist_world() < 2:
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.dev)
        self._reduced_upto = 0

    def _finish_overlapped_reduce(self):
        """End of backward: every live range not yet handed to NCCL goes out now, then the compute stream waits for
        the communication stream.  Leaves the flat buffer holding the SUM over ranks and `grad_scale` = 1 / world."""
        if self._reduced_upto is None:
            return
        import torch.distributed as dist
        self._ready_ranges = []        # whatever is still collected goes out with the remaining ranges below
        done = sorted(self._buckets_pending)
        gaps, cur = [], 0
        for lo, hi in done:
            if lo > cur:
                gaps.append((cur, lo))
            cur = max(cur, hi)
        if cur < self.live_end:
            gaps.append((cur, self.live_end))
        main = torch.cuda.current_stream(self.dev)
        comm = self._comm_stream
        if gaps:
            ready = torch.cuda.Event()
            ready.record(main)
            comm.wait_event(ready)
            with torch.cuda.stream(comm):
                for lo, hi in gaps:
                    if self.time_comm:
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(comm)
                    dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM)
                    if self.time_comm:
                        e1.record(comm)
                        self.comm_events.append(((hi - lo) * 4, e0, e1))
        if self.time_comm:
            self._join_events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self._join_events[0].record(main)
        main.wait_stream(comm)
        if self.time_comm:
            self._join_events[1].record(main)
        self.grad_scale = 1.0 / self._dist_world()
        self._reduced_upto = self.live_end

    def comm_report(self):
        """(after a synchronize) {"allreduce_ms": NCCL time on the communication stream, "exposed_ms": how long the
        compute stream waited for it at the end of backward, "bytes", "buckets"} of the last timed backward."""
        if not self.comm_events:
            return None
        tot = sum(e0.elapsed_time(e1) for _, e0, e1 in self.comm_events)
        exposed = self._join_events[0].elapsed_time(self._join_events[1])
        return {"allreduce_ms": tot, "exposed_ms": exposed, "hidden_ms": max(tot - exposed, 0.0),
                "bytes": sum(b for b, _, _ in self.comm_events), "buckets": len(self.comm_events)}

    # ------------------------------------------------------------------ collective + optimizer
    def all_reduce(self):
        """Gradient all-reduce of the live range of the flat buffer over NCCL / NVLink (reference: DDP,
        base_trainer.py:134-137, which averages).  The buffer is SUMMED in place; the factor that makes it the mean
        (1 / world) is returned and remembered in `self.grad_scale`, which `step()` folds into the clip + Adam kernel,
        so `eng.all_reduce(); eng.step(lr)` trains on the mean gradient like DDP does.  Under torch.distributed the
        backward already calls this (or its overlapped form) before it returns, so that `p.grad` holds the mean as
        DistributedDataParallel leaves it; a second call is a no-op that returns the factor."""
        if self._reduced_upto == self.live_end:     # the backward already all-reduced bucket by bucket (overlapped)
            return self.grad_scale
        from .dp import all_reduce_flat_
        timed = self.time_comm and self._dist_world() > 1
        if timed:
            main = torch.cuda.current_stream(self.dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
        self.grad_scale = all_reduce_flat_(self.flat_grad[:self.live_end])
        if timed:
            e1.record(main)
            self.comm_events = [(int(self.live_end) * 4, e0, e1)]
            self._join_events = (e0, e1)            # one flat call on the compute stream: all of it is exposed
        self._reduced_upto = self.live_end
        return self.grad_scale

    def step(self, lr, lr_scale_text_bert=0.1, lr_scale_mmt=1.0, max_grad_l2_norm=0.25, betas=(0.9, 0.999), eps=1e-8,
             grad_scale=None):
        """clip_grad_norm_(all, max_grad_l2_norm) + Adam.step() over the flat buffers (reference
        base_trainer.py:266-269 with optimizer_attributes of configs/t2s_*.yml).  `grad_scale` multiplies the
        gradients before the clip (default: what the last `all_reduce()` returned, 1 on a single rank)."""
        if grad_scale is None:
            grad_scale = self.grad_scale
        L = _lib.get_lib()
        st = torch.cuda.current_stream(self.dev).cuda_stream
        if not hasattr(self, "_opt_ws"):
            self._opt_ws = torch.empty(1024, device=self.dev, dtype=torch.float64)
            self._sumsq = torch.zeros(1, device=self.dev, dtype=torch.float32)
        self.step_count += 1
        L.sumsq(_ptr(self.flat_grad), self.live_end, _ptr(self._opt_ws), _ptr(self._sumsq), st)
        for gname, scale in (("default", 1.0), ("text_bert", lr_scale_text_bert), ("mmt", lr_scale_mmt)):
            a, b = self.group_ranges[gname]
            if b > a:
                L.adam_step(self.flat_param.data_ptr() + 4 * a, self.flat_grad.data_ptr() + 4 * a,
                            self.adam_m.data_ptr() + 4 * a, self.adam_v.data_ptr() + 4 * a, b - a, lr * scale, betas[0],
                            betas[1], eps, self.step_count, _ptr(self._sumsq), float(max_grad_l2_norm or 0.0), grad_scale, st)
        self._refresh_operand_copies(L, st)      # weights changed: bf16 / hi|lo / transposed operand copies are stale
        return self._sumsq

    # ------------------------------------------------------------------ operand copies after the step (t2s_repack_weights)
    def _refresh_operand_copies(self, L, st):
        """The fused Adam kernel writes the parameters through raw pointers, so every GEMM operand derived from them
        (model._packed, self._wt) is stale after it.  Rebuilding them through torch ops costs ~300 small launches (3.3 ms
        of host time per step); instead one kernel rewrites the existing tensors in place from a job table, built once per
        (packed weights, transposed weights) pair.  Pointers stay put, so cached TMA maps and a captured greedy-decode
        graph keep reading the current weights.  Falls back to dropping the copies when the table cannot be built."""
        m = self.model
        P, W = m._packed, self._wt
        if P is None or W is None or os.environ.get("T2S_B200_REPACK", "1") in ("0", "False"):
            m._packed = None
            self._wt = None
            return
        tab = getattr(self, "_repack_tab", None)
        if tab is None or tab["P"] is not P or tab["W"] is not W:
            tab = self._repack_tab = self._build_repack_table(P, W)
        if tab["jobs"] is None:
            m._packed = None
            self._wt = None
            return
        L.repack_weights(_ptr(self.flat_param), _ptr(tab["jobs"]), tab["n_jobs"], tab["n_tiles"], st)

    def _build_repack_table(self, P, W):
        import struct
        m = self.model
        off, named = self.offsets, self.named
        jobs = []          # (src_off, dst tensor, ld_dst, rows, cols, mode, k_pad)

        def add(src_off, rows, cols, dst, mode, k_pad=0):
            jobs.append((src_off, dst, dst.stride(0) if dst.dim() > 1 else dst.numel(), rows, cols, mode, k_pad))

        def qkv_block(pre):     # q, k, v weights (and biases) are adjacent in the flat buffer (_flatten)
            wq, wk, wv = (pre + "attention.self.%s.weight" % t for t in ("query", "key", "value"))
            bq, bk, bv = (pre + "attention.self.%s.bias" % t for t in ("query", "key", "value"))
            ok = (off[wk] == off[wq] + H * H and off[wv] == off[wk] + H * H and off[bk] == off[bq] + H and off[bv] == off[bk] + H)
            return ok, off[wq], off[bq]

        try:
            for key, prefix, mode in (("text", "text_bert.encoder.layer.%d.", 1), ("qtv", "TransLayer.encoder.layer.%d.", 1),
                                      ("mmt", "mmt.encoder.layer.%d.", 0)):
                for i, lw in enumerate(P.get(key, [])):
                    pre = prefix % i
                    ok, wq_off, bq_off = qkv_block(pre)
                    if not ok:
                        raise LookupError("q/k/v of %s are not adjacent in the flat buffer" % pre)
                    add(wq_off, 3 * H, H, lw["wqkv"], mode, H if mode == 1 else 0)
                    add(bq_off, 1, 3 * H, lw["bqkv"], 3)
                    for name, dst in (("attention.output.dense.weight", "wo"), ("intermediate.dense.weight", "wi"),
                                      ("output.dense.weight", "wo2")):
                        p = named[pre + name]
                        add(off[pre + name], p.shape[0], p.shape[1], lw[dst], mode, p.shape[1] if mode == 1 else 0)
                    wt = W[key][i]
                    add(wq_off, 3 * H, H, wt["wqkvT"], 2)
                    for name, dst in (("attention.output.dense.weight", "woT"), ("intermediate.dense.weight", "wiT"),
                                      ("output.dense.weight", "wo2T")):
                        p = named[pre + name]
                        add(off[pre + name], p.shape[0], p.shape[1], wt[dst], 2)
            for which, lin in (("obj", "linear_obj_feat_to_mmt_in.weight"), ("ocr", "linear_ocr_feat_to_mmt_in.weight")):
                p = named[lin]
                add(off[lin], p.shape[0], p.shape[1], P["w_" + which], 1, P["k_%s_pad" % which])
                add(off[lin], p.shape[0], p.shape[1], W[which + "T"], 2)
            for name, dst, dstT in (("classifier.module.weight", "w_cls", "clsT"), ("ocr_ptr_net.query.weight", "w_ptr_q", "ptr_qT"),
                                    ("ocr_ptr_net.key.weight", "w_ptr_k", "ptr_kT")):
                p = named[name]
                add(off[name], p.shape[0], p.shape[1], P[dst], 0)
                add(off[name], p.shape[0], p.shape[1], W[dstT], 2)
            # everything else in P is an fp32 view of the parameter itself (no copy): check, do not assume
            for n, t in P["f32"].items():
                if n in self.offsets and t.data_ptr() != named[n].data_ptr():
                    raise LookupError("fp32 operand of %s is a copy, not a view" % n)
            for n in self.live_names:
                if not named[n].is_contiguous() or named[n].dtype != torch.float32:
                    raise LookupError("parameter %s is not a contiguous fp32 view" % n)
        except (LookupError, KeyError) as e:
            self.model.writer.write("operand copies are rebuilt through torch after every step (%s)" % e, "warning")
            return dict(P=P, W=W, jobs=None)
        recs, tile0 = [], 0
        for src_off, dst, ld, rows, cols, mode, k_pad in jobs:
            tx = (max(cols, k_pad) + 31) // 32
            ty = (rows + 31) // 32
            recs.append(struct.pack("qqqiiiiii", src_off, dst.data_ptr(), ld, rows, cols, mode, k_pad, tile0, tx))
            tile0 += tx * ty
        blob = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(self.dev)
        return dict(P=P, W=W, jobs=blob, n_jobs=len(recs), n_tiles=tile0, keep=[j[1] for j in jobs])

    # ------------------------------------------------------------------ optimizer state / checkpoints in the reference's format
    def _adam_template(self, config, lr):
        """An (empty-state) torch.optim.Adam over the reference's parameter groups (t2s.py:356-376) -- used only as the
        source of a state_dict layout that `Optimizer.load_state_dict` of this torch accepts."""
        groups = self.model.get_optimizer_parameters(config)
        return groups, torch.optim.Adam(groups, lr=lr, eps=1e-8, weight_decay=0)

    def optimizer_state_dict(self, config, lr=None):
        """The fused Adam state as a `torch.optim.Adam.state_dict()` over `model.get_optimizer_parameters(config)`:
        what the reference's checkpoint stores under "optimizer" (pythia/utils/checkpoint.py:226-232) and restores
        with `optimizer.load_state_dict` (checkpoint.py:114), so a run can move between the two optimizers."""
        lr = float(config.optimizer_attributes.params.lr) if lr is None else lr
        groups, opt = self._adam_template(config, lr)
        sd = opt.state_dict()
        name_of = {id(p): n for n, p in self.named.items()}
        live = set(self.live_names)
        idx = 0
        for g in groups:
            for p in g["params"]:
                n = name_of[id(p)]
                if n in live and self.step_count > 0:
                    o = self.offsets[n]
                    sd["state"][idx] = {"step": torch.tensor(float(self.step_count)),
                                        "exp_avg": self.adam_m[o:o + p.numel()].view_as(p).clone(),
                                        "exp_avg_sq": self.adam_v[o:o + p.numel()].view_as(p).clone()}
                idx += 1
        return sd

    def load_optimizer_state_dict(self, sd, config):
        """Inverse of `optimizer_state_dict` (also accepts the state of a torch.optim.Adam the reference trained)."""
        groups = self.model.get_optimizer_parameters(config)
        name_of = {id(p): n for n, p in self.named.items()}
        self.adam_m.zero_()
        self.adam_v.zero_()
        steps = set()
        idx = 0
        for g in groups:
            for p in g["params"]:
                st = sd["state"].get(idx)
                if st is not None:
                    o = self.offsets[name_of[id(p)]]
                    self.adam_m[o:o + p.numel()].copy_(st["exp_avg"].reshape(-1))
                    self.adam_v[o:o + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.add(int(st["step"]))
                idx += 1
        if len(steps) > 1:
            raise ValueError("the fused Adam keeps one step count; the state has %s" % sorted(steps))
        self.step_count = steps.pop() if steps else 0

    def save_checkpoint(self, path, config, best_iteration=0, best_metric_value=None, lr=None):
        """Same keys as the reference's checkpoint file (checkpoint.py:226-240, minus the git metadata)."""
        torch.save({"model": self.model.state_dict(), "optimizer": self.optimizer_state_dict(config, lr),
                    "best_iteration": best_iteration, "best_metric_value": best_metric_value, "config": config}, path)

    def load_checkpoint(self, path, config):
        """Restore model + optimizer from a checkpoint in the reference's format (checkpoint.py:98-116; a DataParallel
        "module." prefix is stripped like there)."""
        ckpt = torch.load(path, map_location=self.dev, weights_only=False)
        model_sd = ckpt["model"] if "model" in ckpt else ckpt
        model_sd = {(k[7:] if k.startswith("module.") else k): v for k, v in model_sd.items()}
        self.model.load_state_dict(model_sd)           # copies into the views of the flat buffer
        if "optimizer" in ckpt:
            self.load_optimizer_state_dict(ckpt["optimizer"], config)
        self.model._packed = None
        self._wt = None
        self._repack_tab = None
        return ckpt
