"""Host-side mirror of the reference's AudioMamba forward (default configuration) on the B200 engine.

Mirrors /root/reference/src/models/mamba_models.py:193-685 for the configuration every released AuM checkpoint
uses (src/run.py:224-274): rms_norm=True, fused_add_norm=True, residual_in_fp32=True, absolute pos-embed,
one middle cls token, no RoPE, no sequence flips, drop_path 0.  Same constructor argument names and the same
state-dict keys (patch_embed.proj.*, cls_token, pos_embed.pos_embed, layers.N.{mixer.*,norm.weight}, norm_f.weight,
head.*), so reference checkpoints load with strict=True.

This exists because the GPU box has no copy of the reference: bench.py and the whole-model parity tests need a
caller of the hot path.  With the reference tree available, its own src/models/mamba_models.py runs unchanged on
top of the sibling ``mamba_ssm`` package instead.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import mixer, ops
from .modules import Mamba, RMSNorm


class _PatchProj(nn.Module):
    """Holds patch_embed.proj.{weight,bias} with the reference's Conv2d parameter shapes (tokenization.py:224)."""

    def __init__(self, in_chans, embed_dim, patch, device=None, dtype=None):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch, stride=patch, bias=True, device=device, dtype=dtype)
        fan_in = in_chans * patch[0] * patch[1]
        nn.init.trunc_normal_(self.proj.weight, std=(1.0 / fan_in) ** 0.5 / .87962566103423978)
        nn.init.zeros_(self.proj.bias)


class _PosEmbed(nn.Module):
    def __init__(self, n_tokens, embed_dim, device=None, dtype=None):
        super().__init__()
        self.pos_embed = nn.Parameter(torch.zeros(1, n_tokens, embed_dim, device=device, dtype=dtype))
        nn.init.trunc_normal_(self.pos_embed, std=.02)


class _Block(nn.Module):
    def __init__(self, dim, eps, layer_idx, bimamba_type, if_devide_out, device=None, dtype=None):
        super().__init__()
        self.mixer = Mamba(dim, layer_idx=layer_idx, bimamba_type=bimamba_type, if_devide_out=if_devide_out,
                           device=device, dtype=dtype)
        self.norm = RMSNorm(dim, eps=eps, device=device, dtype=dtype)


class AudioMamba(nn.Module):
    def __init__(self, spectrogram_size=(128, 1024), patch_size=(16, 16), strides=(16, 16), depth=24, embed_dim=768,
                 channels=1, num_classes=527, norm_epsilon: float = 1e-5, rms_norm: bool = True,
                 fused_add_norm: bool = True, residual_in_fp32: bool = True, device=None, dtype=None,
                 if_abs_pos_embed=True, if_rope=False, if_cls_token=True, if_bidirectional=False,
                 bimamba_type="v2", if_devide_out=True, use_double_cls_token=False, use_middle_cls_token=True,
                 act_dtype: torch.dtype = torch.float32, use_cuda_graph: bool = False, micro_batches: int = 1,
                 **unsupported):
        super().__init__()
        # The reference's own constructor call (src/run.py:248-274) always passes a set of options that cannot change
        # the function the default path computes - where pretrained weights would come from, how they would be
        # resampled, the pooling used only without a cls token, stochastic depth (identity in eval; 0 in every released
        # recipe).  Those are accepted and ignored; options that DO change the computed function are refused unless they
        # carry the reference's default (src/models/mamba_models.py:194-245).
        ignored = {"imagenet_pretrain_path", "imagenet_pretrain_modelkey", "aum_pretrain_path", "aum_pretrain_fstride",
                   "aum_pretrain_tstride", "pt_hw_seq_len", "imagenet_load_double_cls_token",
                   "imagenet_load_middle_cls_token", "use_PI_for_patch_embed", "final_pool_type", "initializer_cfg",
                   "ft_seq_len", "abs_pos_patch_grid_size", "drop_rate", "if_bimamba", "ssm_cfg"}
        must_be_default = {"imagenet_pretrain": False, "aum_pretrain": False, "bilinear_rope": False,
                           "flip_img_sequences_ratio": -1.0, "transpose_token_sequence": False,
                           "use_end_cls_token": False, "if_rope_residual": False, "flexible_patch_sizes": None,
                           "init_layer_scale": None, "drop_path_rate": 0}
        bad = []
        for k, v in unsupported.items():
            if k in ignored:
                if k == "ssm_cfg" and v:
                    bad.append(k)
                continue
            if k in must_be_default:
                d = must_be_default[k]
                same = (v is None and d is None) or (v is not None and d is not None and type(v) in (bool, int, float)
                                                     and float(v) == float(d))
                if k == "flexible_patch_sizes" and not v:
                    same = True
                if not same:
                    bad.append(k)
                continue
            bad.append(k)
        if bad:
            raise NotImplementedError(f"AudioMamba (B200 mirror): options outside the default AuM path: {sorted(bad)}")
        if not (rms_norm and fused_add_norm and residual_in_fp32 and if_abs_pos_embed and if_cls_token
                and use_middle_cls_token) or if_rope or if_bidirectional or use_double_cls_token or channels != 1:
            raise NotImplementedError("AudioMamba (B200 mirror) implements the default AuM configuration only")
        patch = tuple(patch_size) if isinstance(patch_size, (tuple, list)) else (patch_size, patch_size)
        if tuple(strides) != patch:
            raise NotImplementedError("overlapping patches (strides != patch_size) are not on the default AuM path")
        F_, T_ = spectrogram_size
        self.patch = patch
        self.grid = (F_ // patch[0], T_ // patch[1])
        self.num_patches = self.grid[0] * self.grid[1]
        self.embed_dim = self.d_model = self.num_features = embed_dim
        self.num_classes = num_classes
        self.eps = norm_epsilon
        self.act_dtype = act_dtype
        # Inference fast path: capture the whole forward (all kernel launches of the 24 blocks) in a CUDA graph per
        # input shape and replay it — removes host launch gaps.  Off by default; bench.py turns it on.
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}
        # Inference: split the batch into `micro_batches` independent sequence groups, each on its own CUDA stream,
        # so one group's MUFU-bound scan can overlap another group's tensor-core GEMMs and kernel tails.
        self.micro_batches = micro_batches
        self._streams = None
        self._warm_key = None
        self._grad_sync = None      # (FlatGradReducer, {layer index: chunk}) set by aum_b200.trainer.TrainStep
        fk = {"device": device, "dtype": dtype}
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim, **fk))
        nn.init.trunc_normal_(self.cls_token, std=.02)
        self.head = nn.Linear(embed_dim, num_classes, **fk)
        nn.init.trunc_normal_(self.head.weight, std=.02)
        nn.init.zeros_(self.head.bias)
        self.layers = nn.ModuleList([_Block(embed_dim, norm_epsilon, i, bimamba_type, if_devide_out, **fk)
                                     for i in range(depth)])
        with torch.no_grad():   # _init_weights rescaling of residual projections (mamba_models.py:164-172)
            for blk in self.layers:
                blk.mixer.out_proj.weight.div_(depth ** 0.5)
        self.norm_f = RMSNorm(embed_dim, eps=norm_epsilon, **fk)
        self.patch_embed = _PatchProj(channels, embed_dim, patch, **fk)
        self.pos_embed = _PosEmbed(self.num_patches + 1, embed_dim, **fk)

    # ------------------------------------------------------------------------------------------------
    def _tokens(self, x: torch.Tensor, act: torch.dtype) -> torch.Tensor:
        """(B, T, F) spectrogram -> (B, N+1, Dm) fp32 token sequence with the cls token in the middle and
        the absolute position embedding added (mamba_models.py:510-541, tokenization.py:278-310,414-451)."""
        B, T_, F_ = x.shape
        gf, gt = self.grid
        N = gf * gt
        # im2col of the stride-16 conv (one kernel: gather + cast), the conv as one GEMM, cls / pos assembly (one kernel)
        cols = ops.patchify(x.contiguous(), self.patch, act)
        w = mixer._cache.get(self.patch_embed.proj.weight, f"patchw:{act}",
                             lambda p: p.reshape(p.shape[0], -1).to(act).contiguous())
        tok = ops.gemm_tn(cols, w, bias=mixer._f32(self.patch_embed.proj.bias), out_dtype=torch.float32)
        pos = mixer._cache.get(self.pos_embed.pos_embed, "pos:f32", lambda p: p.float().reshape(N + 1, -1).contiguous())
        cls = mixer._cache.get(self.cls_token, "cls:f32", lambda p: p.float().reshape(-1).contiguous())
        return ops.assemble_tokens(tok.view(B, N, self.embed_dim), pos, cls)

    def _forward_train(self, x: torch.Tensor, return_features: bool) -> torch.Tensor:
        """Autograd path (training): same data flow.  Mixer, add+RMSNorm, the patch-embedding map and the head go through
        the autograd.Functions of aum_b200.autograd (native forward and backward kernels, no cuBLAS / cuDNN); only the
        cls / position assembly (a cat and three adds, once per step) is torch ops."""
        from .autograd import LinearFn
        from .modules import rms_norm_fn
        act = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else self.act_dtype
        B = x.shape[0]
        gf, gt = self.grid
        # stride == kernel conv as im2col (one gather kernel, no gradient needed: x is data) + the engine's linear map,
        # fp32 out as on the inference path
        cols = ops.patchify(x.float().contiguous(), self.patch, act)
        tok = LinearFn.apply(cols, self.patch_embed.proj.weight, self.patch_embed.proj.bias,
                             torch.float32).view(B, gf * gt, self.embed_dim)
        N = tok.shape[1]
        tp = N // 2
        pe = self.pos_embed.pos_embed
        hidden = torch.cat((tok[:, :tp] + pe[:, 1:tp + 1], (self.cls_token + pe[:, :1]).expand(B, -1, -1),
                            tok[:, tp:] + pe[:, tp + 1:]), dim=1)
        residual = None
        sync = self._grad_sync if torch.is_grad_enabled() else None
        for i, blk in enumerate(self.layers):
            if sync is not None and i in sync[1] and hidden.requires_grad:
                # backward reaches this point when every gradient of layers >= i (and of the head) is final: launch
                # that chunk's all-reduce while the earlier layers' backward still runs
                hidden.register_hook(sync[0].hook(sync[1][i]))
            y, residual = rms_norm_fn(hidden, blk.norm.weight, None, residual=residual, prenorm=True,
                                      residual_in_fp32=True, eps=blk.norm.eps)
            hidden = blk.mixer(y.to(act))
        feat = rms_norm_fn(hidden[:, tp, :], self.norm_f.weight, None, residual=residual[:, tp, :], prenorm=False,
                           residual_in_fp32=True, eps=self.norm_f.eps)
        if return_features:
            return feat
        return LinearFn.apply(feat.to(act).contiguous(), self.head.weight, self.head.bias, torch.float32)

    def grad_ready_order(self, n_chunks: int = 3):
        """Parameters in the order their gradients become final during backward (head and final norm first, then layers
        depth-1 .. 0, the embeddings last), the parameter counts after which each all-reduce chunk but the last ends, and
        {layer index: chunk} for the forward hooks: chunk c ends once backward has passed the input of that layer."""
        order = list(self.head.parameters()) + list(self.norm_f.parameters())
        depth = len(self.layers)
        n_chunks = max(1, min(n_chunks, depth))
        # chunk c (c < n_chunks - 1) = head/norm_f (c == 0) + layers [lo_c, hi_c); the last chunk also takes the embeddings
        cuts = [depth - (depth * (c + 1)) // n_chunks for c in range(n_chunks - 1)]      # first layer index of chunks 0..n-2
        chunk_after, hook_layers = [], {}
        hi = depth
        for c, lo in enumerate(cuts):
            for i in range(hi - 1, lo - 1, -1):
                order += list(self.layers[i].parameters())
            chunk_after.append(len(order))
            if lo > 0:
                hook_layers[lo] = c
            hi = lo
        for i in range(hi - 1, -1, -1):
            order += list(self.layers[i].parameters())
        order += [self.pos_embed.pos_embed, self.cls_token] + list(self.patch_embed.parameters())
        assert len(order) == len(list(self.parameters()))
        return order, chunk_after, hook_layers

    def forward_features(self, x: torch.Tensor) -> torch.Tensor:
        L.require_cuda(x)
        act = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else self.act_dtype
        hidden = self._tokens(x.float(), act)
        B, Ntok, Dm = hidden.shape
        tp = (Ntok - 1) // 2
        residual = None
        for blk in self.layers:                                                   # (mamba_models.py:602-622)
            y, residual = ops.add_rmsnorm(hidden.view(B * Ntok, Dm), mixer._f32(blk.norm.weight), None,
                                          residual, eps=blk.norm.eps, prenorm=True, out_dtype=act)   # (:77-97)
            hidden = mixer.mamba_mixer_forward(blk.mixer, y.view(B, Ntok, Dm))                       # (:98)
        # final add+norm (:646-657) is only needed on the cls rows that are returned (:660-664)
        h_cls = hidden[:, tp, :]
        r_cls = residual.view(B, Ntok, Dm)[:, tp, :]
        return ops.add_rmsnorm(h_cls, mixer._f32(self.norm_f.weight), None, r_cls, eps=self.norm_f.eps,
                               prenorm=False, out_dtype=act)

    def _forward_impl(self, x: torch.Tensor, return_features: bool = False) -> torch.Tensor:
        nmb = self.micro_batches
        if nmb > 1 and x.shape[0] % nmb == 0 and x.shape[0] >= 2 * nmb and getattr(self, "serialize_groups", False):
            # same launches, one stream: every kernel runs alone (bench.py's per-kernel timing pass)
            return torch.cat([self._forward_one(xc, return_features) for xc in x.chunk(nmb, dim=0)], dim=0)
        if nmb > 1 and x.shape[0] % nmb == 0 and x.shape[0] >= 2 * nmb:
            # The derived-weight caches (16-bit copies, -exp(A_log), padded dt_proj, ...) are filled by kernels on
            # whichever stream first asks for them; a second stream would then read them with no dependency on those
            # kernels.  So the first forward after construction / load_state_dict / an optimizer step runs its groups
            # one after the other on the current stream (which fills every cache entry there), and only forwards with
            # warm caches fork.
            act = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else self.act_dtype
            warm_key = (self._weights_version(), act, x.device)
            if self._warm_key != warm_key:
                out = torch.cat([self._forward_one(xc, return_features) for xc in x.chunk(nmb, dim=0)], dim=0)
                self._warm_key = warm_key
                return out
            if self._streams is None or len(self._streams) != nmb:
                self._streams = [torch.cuda.Stream(device=x.device) for _ in range(nmb)]
            cur = torch.cuda.current_stream()
            outs = []
            for s_, xc in zip(self._streams, x.chunk(nmb, dim=0)):
                s_.wait_stream(cur)
                with torch.cuda.stream(s_):
                    o = self._forward_one(xc, return_features)
                o.record_stream(cur)
                outs.append(o)
            for s_ in self._streams:
                cur.wait_stream(s_)
            return torch.cat(outs, dim=0)
        return self._forward_one(x, return_features)

    def _forward_one(self, x: torch.Tensor, return_features: bool = False) -> torch.Tensor:
        feat = self.forward_features(x)
        if return_features:
            return feat
        return ops.gemm_tn(feat, mixer._w(self.head.weight, feat.dtype), bias=mixer._f32(self.head.bias),
                           out_dtype=torch.float32)                                                  # (:682)

    def invalidate_graphs(self):
        self._graphs.clear()
        self._warm_key = None

    def _weights_version(self):
        """Changes whenever any parameter is updated in place, reassigned or moved (keys the CUDA graphs and the
        multi-stream warm-cache check)."""
        v = mixer.generation() << 40
        for p_ in self.parameters():
            v += p_._version + (p_.data_ptr() & 0xffff)
        return v

    @torch.no_grad()
    def infer_stream(self, batches, return_features: bool = False):
        """Streaming inference over an iterable of HOST batches (pinned (B, T, F) fp32 tensors of one shape): yields the
        logits of each batch as a device tensor.  The host->device copy of batch i+1 runs on a copy stream while batch
        i's forward (CUDA-graph replay when use_cuda_graph) runs on the current stream, through two device staging
        buffers - so the transfer is off the critical path, which a plain ``model(x_host)`` call per batch cannot do."""
        it = iter(batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        dev = self.cls_token.device
        cur = torch.cuda.current_stream(dev)
        copy_stream = torch.cuda.Stream(device=dev)
        stg = [torch.empty(nxt.shape, device=dev, dtype=nxt.dtype) for _ in range(2)]
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def issue(i, xh):
            copy_stream.wait_event(consumed[i])          # the forward that read this staging buffer has been enqueued and done
            with torch.cuda.stream(copy_stream):
                stg[i].copy_(xh, non_blocking=True)
                copied[i].record(copy_stream)

        for e in consumed:
            e.record(cur)
        issue(0, nxt)
        i = 0
        while nxt is not None:
            try:
                after = next(it)
            except StopIteration:
                after = None
            if after is not None:
                if after.shape != nxt.shape or after.dtype != nxt.dtype:
                    raise ValueError("infer_stream: every batch must have the same shape and dtype")
                issue(i ^ 1, after)
            cur.wait_event(copied[i])
            out = self.forward(stg[i], return_features)   # (graph path: one device-to-device copy into the captured input)
            consumed[i].record(cur)
            yield out
            nxt, i = after, i ^ 1

    def forward(self, x: torch.Tensor, return_features: bool = False) -> torch.Tensor:
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._forward_train(x, return_features)
        if not self.use_cuda_graph or ops.PROFILE is not None:
            return self._forward_impl(x, return_features)
        act = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else self.act_dtype
        wver = self._weights_version()
        dev = self.cls_token.device
        key = (tuple(x.shape), x.dtype, act, return_features)
        ent = self._graphs.get(key)
        if ent is None or ent[3] != wver:
            with torch.no_grad():
                xin = torch.empty(x.shape, device=dev, dtype=x.dtype)   # x may be a (pinned) host tensor
                xin.copy_(x)
                self._forward_impl(xin, return_features)          # eager warm-up: builds the derived-weight caches
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = self._forward_impl(xin, return_features)
            ent = (g, xin, out, wver)
            self._graphs[key] = ent
        g, xin, out, _ = ent
        xin.copy_(x, non_blocking=True)
        g.replay()
        return out.clone()
