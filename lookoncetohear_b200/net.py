"""Drop-in for ``src.models.tfgridnet_realtime.net.Net`` (reference net.py:20-76).

Select it from the reference's config by pointing ``pl_module_args.model`` at
``lookoncetohear_b200.net.Net`` (the plugin boundary is ``utils.import_attr`` at
ts_hear_embed_pl_module.py:25).  Same constructor keywords, same parameter names and shapes (a
Lightning checkpoint's ``state_dict`` loads unchanged), same ``forward / predict /
init_buffers`` signatures.  The arithmetic is NOT here: every call goes through the C ABI of
``liblookonce_b200.so`` (hand-written sm_100a CUDA).  There is no PyTorch/CPU fallback -- on a
machine without the built library or without a CUDA device the calls raise.

The torch modules held below (nn.Conv2d, nn.LSTM, ...) are used purely as parameter containers,
so names / shapes / default initialisation match the reference; their ``forward`` is never run.
"""
import ctypes
import math

import torch
import torch.nn as nn

from . import _cabi

_HOP_DEFAULT = 128


class _Filterbank(nn.Module):
    """Holds the STFT filter matrix as a buffer named like asteroid's ``filterbank._filters``
    (tfgridnet_causal.py:131-135).  Values: sqrt-periodic-Hann windowed, scaled DFT rows
    (Re rows 0..N/2, then Im rows), as asteroid_filterbanks.STFTFB builds them."""

    def __init__(self, n_filters, kernel_size, stride):
        super().__init__()
        n = torch.arange(kernel_size, dtype=torch.float64)
        k = torch.arange(n_filters // 2 + 1, dtype=torch.float64)
        win = torch.sqrt(0.5 - 0.5 * torch.cos(2 * math.pi * n / kernel_size))
        ang = 2 * math.pi * k[:, None] * n[None, :] / n_filters
        scale = 1.0 / (0.5 * math.sqrt(kernel_size * n_filters / stride))
        filt = torch.cat([torch.cos(ang), -torch.sin(ang)], dim=0) * scale
        filt[0] /= math.sqrt(2.0)
        filt[n_filters // 2] /= math.sqrt(2.0)
        self.register_buffer("_filters", (filt * win).unsqueeze(1).float())


class _Codec(nn.Module):
    def __init__(self, n_filters, kernel_size, stride):
        super().__init__()
        self.filterbank = _Filterbank(n_filters, kernel_size, stride)


class _LN(nn.Module):
    """Same names as the reference's LayerNormalization4D / 4DCF wrappers (``.norm.weight``)."""

    def __init__(self, n):
        super().__init__()
        self.norm = nn.LayerNorm(n)


def _attn_branch(emb_dim, out_dim, ln_dim):
    # indices 0 (Linear), 1 (PReLU), 2 (parameter-less reshape), 3 (LayerNorm) as in the reference
    return nn.Sequential(nn.Linear(emb_dim, out_dim), nn.PReLU(), nn.Identity(), _LN(ln_dim))


class _BlockParams(nn.Module):
    """Parameter container for one GridNetBlock (tfgridnet_causal.py:301-401), same creation
    order as the reference so seeded default init reproduces the same values."""

    def __init__(self, emb_dim, n_freqs, hidden, n_head, approx_qk_dim=512):
        super().__init__()
        E = math.ceil(approx_qk_dim * 1.0 / n_freqs)
        self.intra_norm = _LN(emb_dim)
        self.intra_rnn = nn.LSTM(emb_dim, hidden, 1, batch_first=True, bidirectional=True)
        self.intra_linear = nn.Linear(hidden * 2, emb_dim)
        self.inter_norm = _LN(emb_dim)
        self.inter_rnn = nn.LSTM(emb_dim, hidden, 1, batch_first=True, bidirectional=False)
        self.inter_linear = nn.Linear(hidden, emb_dim)
        self.attn_conv_Q = _attn_branch(emb_dim, E * n_head, n_freqs * E)
        self.attn_conv_K = _attn_branch(emb_dim, E * n_head, n_freqs * E)
        self.attn_conv_V = _attn_branch(emb_dim, (emb_dim // n_head) * n_head, n_freqs * (emb_dim // n_head))
        self.attn_concat_proj = _attn_branch(emb_dim, emb_dim, n_freqs * emb_dim)


class _TFGridNetParams(nn.Module):
    def __init__(self, n_fft, stride, spk_emb_dim, emb_dim, n_layers, n_imics, n_srcs, hidden, n_head):
        super().__init__()
        n_freqs = n_fft // 2 + 1
        self.enc = _Codec(n_fft, n_fft, stride)
        self.dec = _Codec(n_fft, n_fft, stride)
        self.conv = nn.Sequential(nn.Conv2d(2 * n_imics, emb_dim, (3, 3), padding=(0, 1)))
        self.blocks = nn.ModuleList([_BlockParams(emb_dim, n_freqs, hidden, n_head) for _ in range(n_layers)])
        self.embed_to_feats_proj = nn.Sequential(nn.Linear(spk_emb_dim, emb_dim * n_freqs),
                                                 nn.LayerNorm(emb_dim * n_freqs))
        self.deconv = nn.ConvTranspose2d(emb_dim, n_srcs * 2, (3, 3), padding=(2, 1))


class SepState(dict):
    """The streaming state returned by ``Net.init_buffers`` and threaded through ``predict``.

    One contiguous device allocation (header + per-stream records, layout in csrc/sep_layout.h,
    exported through ``l2h_sep_state_offsets``) that the kernels update in place; the dict interface
    is kept because reference callers treat the state as an opaque dict they pass back.
    ``to_reference()`` / ``load_reference()`` convert to / from the reference's nested dict of
    tensors (tfgridnet_causal.py:173-186, :408-427).
    """

    _OFFSET_NAMES = ("ring", "k_ld", "k_dim", "v_dim", "att", "st_emb", "st_gate", "st_conv", "st_deconv",
                     "st_istft", "st_blk", "bk_k", "bk_v", "bk_h", "bk_c", "bk_stride")

    def __init__(self, buf, batch, n_blocks, header_bytes, stride, offsets):
        super().__init__()
        self.buf, self.batch, self.n_blocks = buf, batch, n_blocks
        self.header_floats, self.stride = header_bytes // 4, stride
        self.lay = dict(zip(self._OFFSET_NAMES, offsets))
        self["buf"] = buf

    # ---- views ------------------------------------------------------------------------------
    def _rec(self):
        return self.buf[self.header_floats:].view(self.batch, self.stride)

    def header(self):
        """(pos, ncalls) -- synchronises."""
        h = self.buf[:4].view(torch.int64).cpu()
        return int(h[0]), int(h[1])

    def _tails(self, r, par):
        L, B = self.lay, self.batch
        conv = r[:, L["st_conv"]:L["st_deconv"]].view(B, 2, 2, 4, 97)[:, par]            # [B,slot,ch,F]
        deconv = r[:, L["st_deconv"]:L["st_istft"]].view(B, 2, 2, 97, 64)[:, par]        # [B,slot,F,C]
        istft = r[:, L["st_istft"]:L["st_blk"]].view(B, 2, 2, 194)[:, par]               # [B,ear,2F]
        return conv, deconv, istft

    def _block(self, r, i):
        L, B = self.lay, self.batch
        o = L["st_blk"] + i * L["bk_stride"]
        K = r[:, o + L["bk_k"]:o + L["bk_v"]].view(B, 4, L["ring"], L["k_ld"])
        V = r[:, o + L["bk_v"]:o + L["bk_h"]].view(B, 4, L["ring"], L["v_dim"])
        h = r[:, o + L["bk_h"]:o + L["bk_c"]]
        c = r[:, o + L["bk_c"]:o + L["bk_stride"]]
        return K, V, h, c

    def to_reference(self):
        """Nested dict with the reference's keys and shapes (copies; synchronises)."""
        pos, ncalls = self.header()
        L, r, B = self.lay, self._rec(), self.batch
        hist = L["att"] - 1
        conv, deconv, istft = self._tails(r, ncalls & 1)
        out = dict(conv_buf=conv.permute(0, 2, 1, 3).contiguous(),
                   deconv_buf=deconv.permute(0, 3, 1, 2).contiguous(),
                   istft_buf=istft.unsqueeze(-1).contiguous(), gridnet_bufs={})
        # ring slot of frame n is n % ring; history rows are frames pos-49 .. pos-1
        frames = torch.arange(pos - hist, pos)
        slots = torch.remainder(frames, L["ring"]).to(self.buf.device)
        live = (frames >= 0).to(self.buf.device, self.buf.dtype)[None, None, :, None]
        for i in range(self.n_blocks):
            K, V, h, c = self._block(r, i)
            out["gridnet_bufs"][f"buf{i}"] = dict(
                K_buf=(K[:, :, slots, :L["k_dim"]] * live).reshape(B * 4, hist, L["k_dim"]).contiguous(),
                V_buf=(V[:, :, slots] * live).reshape(B * 4, hist, L["v_dim"]).contiguous(),
                h0=h.reshape(1, B * 97, 64).clone(), c0=c.reshape(1, B * 97, 64).clone())
        return out

    def load_reference(self, ref_state):
        """Import a state in the reference's format (the nested dict ``Net.init_buffers`` /
        ``predict`` of the reference produce) so a stream started on the reference implementation can
        be continued here.  The 49 history rows become frames 0..48 of the rings (pos = 49)."""
        L, r, B = self.lay, self._rec(), self.batch
        hist = L["att"] - 1
        dev, dt = self.buf.device, self.buf.dtype
        self.buf.zero_()
        hdr = self.buf[:4].view(torch.int64)
        hdr[0] = hist            # pos: frames consumed so far
        hdr[1] = 0               # ncalls: tails live in parity slot 0
        conv, deconv, istft = self._tails(r, 0)
        conv.copy_(ref_state["conv_buf"].to(dev, dt).permute(0, 2, 1, 3))
        deconv.copy_(ref_state["deconv_buf"].to(dev, dt).permute(0, 2, 3, 1))
        istft.copy_(ref_state["istft_buf"].to(dev, dt)[..., 0])
        for i in range(self.n_blocks):
            K, V, h, c = self._block(r, i)
            g = ref_state["gridnet_bufs"][f"buf{i}"]
            K[:, :, :hist, :L["k_dim"]] = g["K_buf"].to(dev, dt).view(B, 4, hist, L["k_dim"])
            V[:, :, :hist] = g["V_buf"].to(dev, dt).view(B, 4, hist, L["v_dim"])
            h.copy_(g["h0"].to(dev, dt).reshape(B, 97 * 64))
            c.copy_(g["c0"].to(dev, dt).reshape(B, 97 * 64))
        return self


class Net(nn.Module):
    """B200-native replacement of the reference ``Net`` (net.py:20-76)."""

    def __init__(self, stft_chunk_size=160, stft_pad_size=120, embed_dim=256, num_ch=2, D=64, B=6, I=1, J=1,
                 L=0, H=128, use_attn=False, lookahead=True, local_atten_len=100, chunk_causal=False,
                 num_src=2):
        super().__init__()
        self.stft_chunk_size = stft_chunk_size
        self.stft_pad_size = stft_pad_size
        self.num_ch = num_ch
        self.lookahead = lookahead
        self.nfft = stft_chunk_size + stft_pad_size
        self._cfg = _cabi.SepConfig(stft_chunk_size, stft_pad_size, embed_dim, num_ch, D, L, I, J, B, H,
                                    local_atten_len, int(bool(use_attn)), int(bool(lookahead)),
                                    int(bool(chunk_causal)), num_src)
        self.num_src, self.embed_dim, self.n_blocks = num_src, embed_dim, B
        self.tfgridnet = _TFGridNetParams(self.nfft, stft_chunk_size, embed_dim, D, B, num_ch, num_src, H,
                                          max(L, 1))
        self._handle = None
        self._dirty = True                      # weights need (re)packing into the engine
        self._ws = None
        # batch*frames per kernel chain.  36 clips of 500 frames: the intra recurrence (2 x 18 000 sequences, 32 per tensor-core CTA) is
        # 3.8 waves of the 296 resident CTAs and the inter recurrence (3 492 sequences, 4 per CUDA-core CTA) 2.95 waves; measured
        # 651 k frames/s against 628 k at 18 clips, 600 k at 16 (tools/offline_split_experiment.py)
        self.max_frames_per_launch = 18432

    # ---- engine plumbing -------------------------------------------------------------------
    def _engine(self):
        if self._handle is None:
            h = ctypes.c_void_p()
            _cabi.check(_cabi.lib().l2h_sep_create(ctypes.byref(self._cfg), ctypes.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None:
                _cabi.lib().l2h_sep_destroy(self._handle)
        except Exception:
            pass

    def refresh_weights(self):
        """Call after editing parameters in place; load_state_dict / .to() / .cuda() do it themselves."""
        self._dirty = True

    def __getstate__(self):
        # copy.deepcopy / pickle: the native handle, workspace and staging buffers are per-instance
        d = dict(self.__dict__)
        d["_handle"], d["_ws"], d["_dirty"] = None, None, True
        d.pop("_host_stage", None)
        d.pop("_predict_stage", None)
        d.pop("_last_stream_state", None)
        d["_cfg"] = bytes(self._cfg)
        return d

    def __setstate__(self, d):
        d = dict(d)
        d["_cfg"] = _cabi.SepConfig.from_buffer_copy(d["_cfg"])
        self.__dict__.update(d)

    def _apply(self, fn, *args, **kwargs):
        self._dirty = True
        return super()._apply(fn, *args, **kwargs)

    def _sync_weights(self, device):
        """(Re)pack the weights into the engine after construction, load_state_dict or .to()."""
        if not self._dirty:
            return
        tensors = dict(self.state_dict())
        h, L = self._engine(), _cabi.lib()
        for name, t in tensors.items():
            if name.endswith("torch_window"):
                continue
            host = t.detach().to("cpu", torch.float32).contiguous()
            _cabi.check(L.l2h_sep_load_weight(h, name.encode(), host.data_ptr(), host.numel()))
        with torch.cuda.device(device):
            _cabi.check(L.l2h_sep_commit_weights(h, torch.cuda.current_stream(device).cuda_stream))
        self._dirty = False

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # real asteroid registers an extra (redundant) window buffer on the filterbank; accept it
        for k in [k for k in state_dict if k.startswith(prefix) and k.endswith("filterbank.torch_window")]:
            state_dict.pop(k)
        self._dirty = True
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def _workspace(self, device, batch, frames, flags=0):
        n = ctypes.c_size_t()
        _cabi.check(_cabi.lib().l2h_sep_workspace_bytes(self._engine(), batch, frames, flags, ctypes.byref(n)))
        if self._ws is None or self._ws.numel() < n.value or self._ws.device != device:
            self._ws = torch.empty(n.value, dtype=torch.uint8, device=device)
        return self._ws, n.value

    def _stream_workspace(self, device, batch, chunks_per_call):
        n = ctypes.c_size_t()
        _cabi.check(_cabi.lib().l2h_sep_stream_workspace_bytes(self._engine(), batch, chunks_per_call, ctypes.byref(n)))
        if self._ws is None or self._ws.numel() < n.value or self._ws.device != device:
            self._ws = torch.empty(n.value, dtype=torch.uint8, device=device)
        return self._ws, n.value

    def set_option(self, name, value):
        """Engine switches: "pipeline" (wavefront pipelining of one-hop calls), "pdl", "fused_mid", "pipeline_frames",
        "pipeline_split_mid" and the lanes per pipeline stage: "pipeline_lanes" (BiLSTM), "pipeline_qkv_lanes",
        "pipeline_attn_lanes", "pipeline_out_lanes", "pipeline_front_lanes", "pipeline_back_lanes"."""
        _cabi.check(_cabi.lib().l2h_sep_set_option(self._engine(), name.encode(), int(value)))

    def reset_options(self):
        """Pipeline lane counts / mid split / hops per graph back to the engine defaults."""
        self.set_option("defaults", 0)

    def pipeline_frames(self):
        k = ctypes.c_int32()
        _cabi.check(_cabi.lib().l2h_sep_pipeline_frames(self._engine(), ctypes.byref(k)))
        return k.value

    @staticmethod
    def _require_cuda(t):
        if not t.is_cuda:
            raise RuntimeError("lookoncetohear_b200.Net runs only on a CUDA (sm_100a) device: the hot path is "
                               "hand-written CUDA with no CPU fallback")

    # ---- reference API ---------------------------------------------------------------------
    def init_buffers(self, batch_size, device, out=None):
        """Fresh streaming state for `batch_size` streams (net.py:40-41 of the reference).  `out`: a SepState of the same batch
        size to re-initialise IN PLACE -- the engine's cached CUDA graphs are keyed on the state's address, so a service that
        resets a stream keeps its graphs warm this way (a new allocation means one more capture + instantiation of the
        clip-sized pipelined graph, ~0.1-0.4 s)."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("lookoncetohear_b200.Net.init_buffers: CUDA device required (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        L, h = _cabi.lib(), self._engine()
        n = ctypes.c_size_t()
        _cabi.check(L.l2h_sep_state_bytes(h, batch_size, ctypes.byref(n)))
        hb, stride, offs = self._state_layout()
        if out is not None and (out.batch != batch_size or out.buf.device != device or out.buf.numel() != n.value // 4):
            raise ValueError("init_buffers(out=...): the state to reuse has another batch size, device or layout")
        buf = out.buf if out is not None else torch.empty(n.value // 4, dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _cabi.check(L.l2h_sep_state_init(h, buf.data_ptr(), batch_size,
                                             torch.cuda.current_stream(device).cuda_stream))
        return SepState(buf, batch_size, self.n_blocks, hb, stride, offs)

    def _state_layout(self):
        """(header bytes, floats per stream record, record offsets) from the C side."""
        L, h = _cabi.lib(), self._engine()
        hb, stride = ctypes.c_int64(), ctypes.c_int64()
        _cabi.check(L.l2h_sep_state_layout(h, ctypes.byref(hb), ctypes.byref(stride)))
        offs = (ctypes.c_int64 * 16)()
        _cabi.check(L.l2h_sep_state_offsets(h, offs, 16))
        return hb.value, stride.value, list(offs)

    def _run(self, x, embed, state, frames, out_len, flags=0):
        """x [B,M,n] (any length; samples beyond n read as zero), embed [B,256]."""
        self._require_cuda(x)
        dev = x.device
        self._sync_weights(dev)
        x = x.contiguous().float()
        embed = embed.to(dev, torch.float32).contiguous()
        Bsz = x.shape[0]
        if state.batch != Bsz:
            raise ValueError(f"state was built for batch {state.batch}, input has batch {Bsz}")
        y = torch.empty(Bsz, self.num_src, out_len, dtype=torch.float32, device=dev)
        ws, nbytes = self._workspace(dev, Bsz, frames, flags)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().l2h_sep_forward(
                self._engine(), x.data_ptr(), x.stride(0), x.stride(1), x.shape[-1], embed.data_ptr(),
                state.buf.data_ptr(), y.data_ptr(), y.stride(0), y.stride(1), out_len, Bsz, frames,
                ws.data_ptr(), ws.numel(), flags, torch.cuda.current_stream(dev).cuda_stream))
        return y

    def predict(self, x, embed, input_state, pad=True):
        """Reference net.py:54-66.  x [B,M,N]; embed [B,256]; returns (y [B,S,*], state)."""
        hop, la = self.stft_chunk_size, self.stft_pad_size
        n = x.shape[-1]
        if pad:
            frames = (n + hop - 1) // hop          # mod-pad to whole hops, + look-ahead zeros
            out_len = n
        else:
            if (n - la) % hop != 0 or n < hop + la:
                raise ValueError(f"pad=False needs {hop}*T+{la} samples, got {n}")
            frames = (n - la) // hop
            out_len = frames * hop
        if not isinstance(input_state, SepState):
            raise TypeError("input_state must come from Net.init_buffers()")
        y = self._run(x, embed, input_state, frames, out_len)
        return y, input_state

    def forward(self, x, embeds, input_state=None, pad=True):
        """Reference net.py:68-76.  x [B,M,N]; embeds [B,1,256] -> [B,S,N]."""
        embeds = embeds[:, 0]
        Bsz = x.shape[0]
        hop = self.stft_chunk_size
        frames = (x.shape[-1] + hop - 1) // hop
        # independent streams: split the batch so batch*frames stays inside the workspace bound
        per = max(1, self.max_frames_per_launch // max(frames, 1))
        if input_state is not None or Bsz <= per:
            if input_state is None:
                input_state = self.init_buffers(Bsz, x.device)
            y, _ = self.predict(x, embeds, input_state, pad)
            return y
        outs = []
        for b0 in range(0, Bsz, per):
            xs, es = x[b0:b0 + per], embeds[b0:b0 + per]
            st = self.init_buffers(xs.shape[0], x.device)
            outs.append(self.predict(xs, es, st, pad)[0])
        return torch.cat(outs, dim=0)

    def stream_dev(self, x_dev, embed_dev, chunks_per_call=1, state=None, n_calls=None, out=None):
        """Streaming over a device-resident clip (l2h_sep_stream_dev): x_dev [B,M,N] is consumed
        chunks_per_call hops per call with the state carried, every call one CUDA-graph replay.
        Returns y [B,S,N] (device).  Asynchronous."""
        self._require_cuda(x_dev)
        dev = x_dev.device
        self._sync_weights(dev)
        hop = self.stft_chunk_size
        x = x_dev.contiguous().float()
        Bsz, _, n = x.shape
        step = hop * chunks_per_call
        if n_calls is None:
            n_calls = (n + step - 1) // step
        if state is None:
            state = self.init_buffers(Bsz, dev)
        y = out if out is not None else torch.empty(Bsz, self.num_src, n, dtype=torch.float32, device=dev)
        ws, _ = self._stream_workspace(dev, Bsz, chunks_per_call)
        emb = embed_dev.to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().l2h_sep_stream_dev(
                self._engine(), x.data_ptr(), n, emb.data_ptr(), state.buf.data_ptr(), y.data_ptr(), n, Bsz,
                n_calls, chunks_per_call, ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        self._last_stream_state = state
        return y

    def stream_host(self, x_host, embed_dev, chunks_per_call=1, state=None, out=None):
        """End-to-end streaming with HOST buffers (l2h_sep_stream_host): x_host [B,M,N] CPU tensor
        (pinned here if it is not); every round copies its samples host->device, runs the one-hop /
        multi-hop chains and copies the new samples back (a round = one call of chunks_per_call hops, or
        with chunks_per_call == 1 a pipelined group of up to pipeline_frames() hops).  Returns
        y [B,S,N] on the host (`out`: a pinned [B,S,>=N] tensor to write into).  Synchronises."""
        dev = embed_dev.device
        self._require_cuda(embed_dev)
        self._sync_weights(dev)
        hop, la = self.stft_chunk_size, self.stft_pad_size
        Bsz, _, n = x_host.shape
        step = hop * chunks_per_call
        n_calls = (n + step - 1) // step
        xh = x_host.contiguous().float()
        if not xh.is_pinned():
            xh = xh.pin_memory()
        grp = max(chunks_per_call, self.pipeline_frames() if chunks_per_call == 1 else 1)
        key = (Bsz, n_calls * step, grp, str(dev))
        cache = getattr(self, "_host_stage", None)
        if cache is None or cache[0] != key:        # staging buffers are reused across calls
            yh_c = torch.empty(Bsz, self.num_src, n_calls * step, dtype=torch.float32).pin_memory()
            xs = torch.empty(Bsz, self.num_ch, hop * grp + la, dtype=torch.float32, device=dev)
            ys = torch.empty(Bsz, self.num_src, hop * grp, dtype=torch.float32, device=dev)
            cache = (key, yh_c, xs, ys)
            self._host_stage = cache
        _, yh_c, xs, ys = cache
        yh = out if out is not None else yh_c
        if not yh.is_pinned() or yh.shape[-1] < n or not yh.is_contiguous():
            raise ValueError("out must be a contiguous pinned [B, S, >= N] float32 tensor")
        if state is None:
            state = self.init_buffers(Bsz, dev)
        ws, _ = self._stream_workspace(dev, Bsz, chunks_per_call)
        emb = embed_dev.to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().l2h_sep_stream_host(
                self._engine(), xh.data_ptr(), n, emb.data_ptr(), state.buf.data_ptr(), yh.data_ptr(),
                yh.shape[-1], Bsz, n_calls, chunks_per_call, xs.data_ptr(), ys.data_ptr(), ws.data_ptr(),
                ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        self._last_stream_state = state
        return yh[..., :n] if out is not None else yh[..., :n].clone()

    def predict_host(self, x_host, embed_dev, state, out=None):
        """One streaming call with HOST buffers, natively (l2h_sep_stream_host with one call): x_host [B,M,128*T+64] PINNED (the
        chunk plus the 64 look-ahead samples, as predict(..., pad=False) takes it) is copied host->device, the chain runs, the
        128*T new samples per ear are copied back into `out` ([B,S,128*T] pinned; allocated once if None) and the stream is
        synchronised -- the copies, the launch and the wait are one C call instead of four Python-level ops.  Returns (out, state)."""
        dev = embed_dev.device
        self._require_cuda(embed_dev)
        self._sync_weights(dev)
        hop, la = self.stft_chunk_size, self.stft_pad_size
        Bsz, _, n = x_host.shape
        if (n - la) % hop != 0 or n < hop + la:
            raise ValueError(f"predict_host needs {hop}*T+{la} samples, got {n}")
        frames = (n - la) // hop
        if not (x_host.is_pinned() and x_host.is_contiguous() and x_host.dtype == torch.float32):
            raise ValueError("x_host must be a contiguous pinned float32 tensor")
        if not isinstance(state, SepState):
            raise TypeError("state must come from Net.init_buffers()")
        key = (Bsz, frames, str(dev))
        cache = getattr(self, "_predict_stage", None)
        if cache is None or cache[0] != key:
            cache = (key, torch.empty(Bsz, self.num_src, hop * frames, dtype=torch.float32).pin_memory(),
                     torch.empty(Bsz, self.num_ch, n, dtype=torch.float32, device=dev),
                     torch.empty(Bsz, self.num_src, hop * frames, dtype=torch.float32, device=dev),
                     self._stream_workspace(dev, Bsz, frames)[0])
            self._predict_stage = cache
        _, yh_c, xs, ys, ws = cache
        emb = embed_dev if (embed_dev.dtype == torch.float32 and embed_dev.is_contiguous()) else embed_dev.to(torch.float32).contiguous()
        yh = out if out is not None else yh_c
        if not yh.is_pinned() or not yh.is_contiguous() or tuple(yh.shape) != (Bsz, self.num_src, hop * frames):
            raise ValueError("out must be a contiguous pinned [B, S, 128*T] float32 tensor")
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().l2h_sep_stream_host(
                self._engine(), x_host.data_ptr(), n, emb.data_ptr(), state.buf.data_ptr(), yh.data_ptr(), hop * frames, Bsz, 1,
                frames, xs.data_ptr(), ys.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        return yh, state

    # ---- debugging aid for the parity tests ------------------------------------------------------
    def forward_with_taps(self, x, embeds):
        """Whole-utterance forward that also returns the activations after every stage
        ([B,T,97,64] each): encoder, then per block (after intra, after inter, block output)."""
        embeds = embeds[:, 0]
        hop = self.stft_chunk_size
        frames = (x.shape[-1] + hop - 1) // hop
        st = self.init_buffers(x.shape[0], x.device)
        y = self._run(x, embeds, st, frames, x.shape[-1], flags=1)
        off, ns = ctypes.c_int64(), ctypes.c_int32()
        _cabi.check(_cabi.lib().l2h_sep_tap_info(self._engine(), x.shape[0], frames, ctypes.byref(off),
                                                ctypes.byref(ns)))
        n = x.shape[0] * frames * 97 * 64
        wsf = self._ws.view(torch.float32)
        taps = [wsf[off.value + i * n: off.value + (i + 1) * n].view(x.shape[0], frames, 97, 64).clone()
                for i in range(ns.value)]
        return y, taps, st
