"""Host-side engine of K3: runs the reference's MLPs as chains of tcgen05 GEMM launches.

The nn.Modules in models/ keep the reference's fp32 parameters (names, [out,in] shapes); this
file owns the cached bf16 operand copies (zero-padded W and W^T per layer), the launch
sequences for forward / backward and the autograd glue.  Dense layers run in
nfs_linear_bf16 (forward and dgrad, with bias / ReLU / sigmoid / ReLU-mask epilogues) and
nfs_wgrad_bf16 (weight + bias gradients); nothing here computes on the CPU.
"""
import ctypes
import os

import torch

from . import _lib, ops
from ._lib import ptr
from .ops import _stream


_WEIGHT_EPOCH = 0
_SIDE = {}


def _side_stream(dev):
    """One auxiliary stream per device for launches that may overlap the current stream's."""
    key = str(dev)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


def bump_weight_epoch():
    """Called by optimisers that update parameter storage behind autograd's back (FusedAdam):
    invalidates every cached bf16 operand copy."""
    global _WEIGHT_EPOCH
    _WEIGHT_EPOCH += 1


def _ceil_to(v, m):
    return (v + m - 1) // m * m


def pad_in(k):
    """K (reduction) padding of a GEMM operand: multiple of 64 (one 128-byte swizzle row)."""
    return _ceil_to(k, 64)


def pad_hidden(n):
    """Hidden widths are both an N (<= 256, % 32) and the next layer's K (% 64), and the wgrad
    kernel wants its M side in {128, 256}."""
    if n <= 128:
        return 128
    if n <= 256:
        return 256
    raise RuntimeError("nfs_b200: hidden widths above 256 are not supported by the tcgen05 dense kernels")


class PackedLinear:
    """bf16 operand copies of one (or several row-stacked) fp32 nn.Linear layers."""

    def __init__(self, linears, k_pad, n_pad):
        self.linears = list(linears)
        self.k_pad, self.n_pad = k_pad, n_pad
        self.k = self.linears[0].in_features
        self.rows = [l.out_features for l in self.linears]
        self.n = sum(self.rows)
        assert self.n <= n_pad and self.k <= k_pad
        self._key = None
        self.w16 = self.w16t = self.bias = None

    def current_key(self):
        params = [p for l in self.linears for p in (l.weight, l.bias)]
        return (str(params[0].device), _WEIGHT_EPOCH) + tuple((p.data_ptr(), p._version) for p in params)

    def alloc(self, dev):
        if self.w16 is None or self.w16.device != dev:
            self.w16 = torch.zeros(self.n_pad, self.k_pad, device=dev, dtype=torch.bfloat16)
            self.w16t = torch.zeros(self.k_pad, self.n_pad, device=dev, dtype=torch.bfloat16)
            self.bias = torch.zeros(self.n_pad, device=dev, dtype=torch.float32)

    def table_rows(self):
        """Rows of an nfs_pack_table table that refresh this layer's operand copies (after alloc())."""
        rows, row = [], 0
        for l in self.linears:
            w, b, n, k = l.weight, l.bias, l.out_features, l.in_features
            rows.append([w.data_ptr(), n, k, k, self.w16.data_ptr(), self.k_pad, row, 0, 0, 0])
            rows.append([w.data_ptr(), n, k, k, self.w16t.data_ptr(), self.n_pad, 0, row, 1, 0])
            rows.append([b.data_ptr(), n, 0, 0, self.bias.data_ptr(), 0, row, 0, 2, 0])
            row += n
        return rows

    def refresh(self):
        params = [p for l in self.linears for p in (l.weight, l.bias)]
        dev = params[0].device
        key = (str(dev), _WEIGHT_EPOCH) + tuple((p.data_ptr(), p._version) for p in params)
        if key == self._key:
            return self
        if not params[0].is_cuda:
            raise RuntimeError("nfs_b200: model parameters must live on a CUDA device (no CPU fallback)")
        self.alloc(dev)
        row = 0
        with torch.cuda.device(dev), torch.no_grad():
            for l in self.linears:
                w = l.weight.detach()
                if w.dtype != torch.float32 or not w.is_contiguous():
                    w = w.float().contiguous()
                _lib.call("nfs_pack_linear_bf16", ptr(w), l.out_features, l.in_features, self.n_pad, self.k_pad,
                          row, 0, ptr(self.w16), ptr(self.w16t), _stream())
                self.bias[row:row + l.out_features].copy_(l.bias.detach())
                row += l.out_features
        self._key = key
        return self


def bias_terms(b):
    """fp32 bias vector -> [n, 8] bf16 operand of the chain kernel's bias MMA (nfs_bias_terms_bf16)."""
    out = torch.empty((b.numel(), 8), device=b.device, dtype=torch.bfloat16)
    with torch.cuda.device(b.device):
        _lib.call("nfs_bias_terms_bf16", ptr(b), b.numel(), ptr(out), _stream())
    return out


_freq_cache = {}
_pow2_cache = {}


def first_band(freqs):
    """freqs[0] as a Python float (cached with the octave test: one read-back for a table on the GPU)."""
    bands_are_octaves(freqs)
    return _pow2_cache[(freqs.data_ptr(), freqs._version, int(freqs.numel()), str(freqs.device), "f0")]


def bands_are_octaves(freqs):
    """True when freqs[k] == freqs[0] * 2^k exactly (decided once per table on the host; a table that lives
    on the GPU is read back once)."""
    key = (freqs.data_ptr(), freqs._version, int(freqs.numel()), str(freqs.device))
    hit = _pow2_cache.get(key)
    if hit is None:
        f = freqs.detach().double().cpu()
        hit = bool(f.numel() > 1 and torch.equal(f, f[0] * 2.0 ** torch.arange(f.numel(), dtype=torch.float64)))
        if len(_pow2_cache) > 64:
            _pow2_cache.clear()
        _pow2_cache[key] = hit
        _pow2_cache[key + ("f0",)] = float(f[0]) if f.numel() else 1.0
    return hit


def freqs_on(device, freqs):
    """fp32 copy of a frequency table on `device`, cached (the reference keeps freq_bands on the CPU,
    positional_encoding.py:14-18; copying it per call would put a pageable H2D copy in every step)."""
    if freqs.device == device and freqs.dtype == torch.float32 and freqs.is_contiguous():
        return freqs.detach()
    key = (freqs.data_ptr(), freqs._version, int(freqs.numel()), str(device))
    hit = _freq_cache.get(key)
    if hit is None:
        if len(_freq_cache) > 64:
            _freq_cache.clear()
        hit = _freq_cache[key] = freqs.detach().to(device=device, dtype=torch.float32).contiguous()
    return hit


def encode_operand(x, freqs, k_pad, extra=None, gate=None, out=None):
    """fp32 rows -> bf16 GEMM operand [P,k_pad] = [enc(x) * g0 | extra * g1 | 0] (nfs_posenc_bf16).
    freqs=None copies x itself (already-encoded input); gate (P,2) fp32 = the softmax gate of
    dino_feature_model.py:188-192 (None = 1).  `out`: a bf16 [P,k_pad] column block of a wider tensor."""
    x = ops._f32c(x)
    P, D = x.shape
    L = 0 if freqs is None else int(freqs.numel())
    E = 0 if extra is None else extra.shape[-1]
    pitch = 0
    if out is None:
        out = torch.empty((P, k_pad), device=x.device, dtype=torch.bfloat16)
    else:
        if out.dtype != torch.bfloat16 or out.shape != (P, k_pad) or out.stride(1) != 1:
            raise RuntimeError("encode_operand: out must be a bf16 [P,k_pad] block with contiguous columns")
        pitch = out.stride(0)
    if P:
        fr = None if freqs is None else freqs_on(x.device, freqs)
        octaves = freqs is not None and bands_are_octaves(freqs)
        g0 = g1 = None
        if gate is not None:
            gate = ops._f32c(gate)
            g0, g1 = ptr(gate), ctypes.c_void_p(gate.data_ptr() + 4)
        extra = ops._f32c(extra) if E else None          # a converted copy must outlive the launch: hold it in a name
        with torch.cuda.device(x.device):
            _lib.call("nfs_posenc_bf16", ptr(x), ptr(fr), ptr(extra), g0, g1, 2, P, D, L, E,
                      k_pad, int(pitch), int(octaves), ptr(out), _stream())
    return out


def act_grad(out, g_out, act, n_pad, dst=None):
    """dY bf16 [P,n_pad] = g_out * act'(out) (nfs_act_grad_bf16); `dst`: column block of a wider tensor."""
    out, g_out = ops._f32c(out), ops._f32c(g_out)
    P, C = out.shape
    pitch = 0
    if dst is None:
        dst = torch.empty((P, n_pad), device=out.device, dtype=torch.bfloat16)
    else:
        pitch = dst.stride(0)
    if P:
        with torch.cuda.device(out.device):
            _lib.call("nfs_act_grad_bf16", ptr(out), ptr(g_out), P, C, int(act), n_pad, int(pitch), ptr(dst), _stream())
    return dst


# --------------------------------------------------------------------------------- G1
class G1Plan:
    """Launch plan for nerf_model.NeRFMLP (nerf_model.py:5-24): n_layers x (Linear + ReLU),
    then the [rgb|sigma] head = rgb_out and sigma_out stacked into one 64-row operand."""

    def __init__(self, module):
        self.module = module
        layers = list(module.layers)
        self.in_dim = layers[0].in_features
        self.hidden = layers[0].out_features
        self.k0 = pad_in(self.in_dim)
        self.h_pad = pad_hidden(self.hidden)
        self.packed = [PackedLinear([l], self.k0 if i == 0 else self.h_pad, self.h_pad) for i, l in enumerate(layers)]
        self.head = PackedLinear([module.rgb_out, module.sigma_out], self.h_pad, 64)

    def params(self):
        m = self.module
        ps = []
        for l in m.layers:
            ps += [l.weight, l.bias]
        ps += [m.sigma_out.weight, m.sigma_out.bias, m.rgb_out.weight, m.rgb_out.bias]
        return ps

    def refresh(self):
        """Bring the bf16 operands up to date with the fp32 parameters.  On the fused path that is ONE launch
        (nfs_pack_stack writes every layer's W, W^T and bias terms into the stacked operands); the per-layer copies
        of the layer-by-layer route are refreshed lazily (refresh_layers)."""
        params = self.params()
        dev = params[0].device
        if not params[0].is_cuda:
            raise RuntimeError("nfs_b200: model parameters must live on a CUDA device (no CPU fallback)")
        key = (str(dev), _WEIGHT_EPOCH) + tuple((p.data_ptr(), p._version) for p in params)
        if key == getattr(self, "_key", None):
            return
        place = (str(dev),) + tuple((p.data_ptr(), p.dtype, p.is_contiguous()) for p in params)
        if place != getattr(self, "_place", None):
            self._build_stack(dev)
            self._place = place
        self._layers_stale = True
        if self.fusable:
            if self._table is None or os.environ.get("NFS_PACK_STACK", "1") == "0":   # non-contiguous / non-fp32 parameters (or the developer switch): through the layer copies
                self.refresh_layers()
                self._copy_layers_into_stack()
            else:
                with torch.cuda.device(dev), torch.no_grad():
                    _lib.call("nfs_pack_stack", ptr(self._table), self._table.shape[0], self._max_elems,
                              ptr(self.w_stack), ptr(self.wt_stack), ptr(self.b_stack), _stream())
        else:
            self.refresh_layers()
        self._key = key

    def refresh_layers(self):
        """Per-layer operand copies (PackedLinear) used by the layer-by-layer route."""
        if getattr(self, "_layers_stale", True):
            for p in self.packed:
                p.refresh()
            self.head.refresh()
            self._layers_stale = False

    def _build_stack(self, dev):
        """Allocate the stacked operands of the fused chain - one [rows,256] bf16 tensor holding every layer's padded
        W, one holding the W^T of the dgrad chain, the [rows,8] bias terms - and the table nfs_pack_stack fills
        them from."""
        layers = self.packed + [self.head]
        self.fusable = self.k0 <= 256 and len(layers) <= 12
        if not self.fusable:
            return
        rows = sum(p.n_pad for p in layers)
        r, row0 = 0, []
        for p in layers:
            row0.append(r)
            r += p.n_pad
        self.w_stack = torch.zeros(rows, 256, device=dev, dtype=torch.bfloat16)
        self.b_stack = torch.zeros(rows, 8, device=dev, dtype=torch.bfloat16)
        self.w_rows = rows
        n = len(layers)
        arr = ctypes.c_int32 * n
        self.c_k = arr(*[p.k_pad for p in layers])
        self.c_n = arr(*[p.n_pad for p in layers])
        self.c_act = arr(*([1] * (n - 1) + [2]))
        self.c_row0 = arr(*row0)
        # dgrad chain: X = dL/d(head pre-activation) [P,64]; layer j multiplies by the transposed
        # weight of forward layer (n_hidden - j) (j = 0: the head) and masks with that layer's input
        nh = n - 1
        back = [self.head] + [self.packed[i] for i in range(nh - 1, 0, -1)]
        rows_t = sum(p.k_pad for p in back)
        r, row0_t = 0, {}
        for p in back:
            row0_t[id(p)] = r
            r += p.k_pad
        self.wt_stack, self.wt_rows = torch.zeros(rows_t, 256, device=dev, dtype=torch.bfloat16), rows_t
        arr_b = ctypes.c_int32 * nh
        self.cb_k = arr_b(*[p.n_pad for p in back])
        self.cb_n = arr_b(*[p.k_pad for p in back])
        self.cb_act = arr_b(*([4] * nh))
        self.cb_row0 = arr_b(*[row0_t[id(p)] for p in back])
        self.cb_mask = arr_b(*[nh - 1 - j for j in range(nh)])      # h_{nh-j} = forward save[nh-1-j]
        # nfs_pack_stack table: one row per parameter tensor
        self._row0, self._row0_t = row0, row0_t
        table, ok = [], True
        for p, r0 in zip(layers, row0):
            sub = 0
            for lin in p.linears:
                for t in (lin.weight, lin.bias):
                    ok = ok and t.dtype == torch.float32 and t.is_contiguous()
                table.append([lin.weight.data_ptr(), lin.out_features, lin.in_features, r0 + sub, 0,
                              row0_t.get(id(p), -1), sub, 0])
                table.append([lin.bias.data_ptr(), lin.out_features, 0, 0, 0, -1, 0, r0 + sub])
                sub += lin.out_features
        self._table = torch.tensor(table, dtype=torch.int64, device=dev) if ok else None
        self._max_elems = max(row[1] * max(row[2], 1) for row in table)

    def _copy_layers_into_stack(self):
        layers = self.packed + [self.head]
        b = torch.zeros(self.w_rows, device=self.w_stack.device, dtype=torch.float32)
        for p, r in zip(layers, self._row0):
            self.w_stack[r:r + p.n_pad, :p.k_pad].copy_(p.w16)
            b[r:r + p.n_pad].copy_(p.bias)
            if id(p) in self._row0_t:
                rt = self._row0_t[id(p)]
                self.wt_stack[rt:rt + p.k_pad, :p.n_pad].copy_(p.w16t)
        self.b_stack.copy_(bias_terms(b))

    def can_encode_in_kernel(self, freqs, dim=3):
        """The chain kernel can build the first layer's operand itself (positions -> sin/cos encoding)."""
        return (getattr(self, "fusable", False) and self.k0 == 64 and dim == 3 and 1 <= int(freqs.numel()) <= 10
                and self.in_dim == 3 * (2 * int(freqs.numel()) + 1) and bands_are_octaves(freqs)
                and os.environ.get("NFS_MLP_FUSED", "1") != "0" and os.environ.get("NFS_MLP_FUSED_ENC", "1") != "0")

    def chain_model(self, freqs):
        """struct nfs_chain_model for nfs_render_fused_fwd (host struct; the arrays it points to live in the plan)."""
        return _lib.ChainModel(len(self.packed) + 1, ctypes.cast(self.c_k, ctypes.c_void_p), ctypes.cast(self.c_n, ctypes.c_void_p),
                               ctypes.cast(self.c_act, ctypes.c_void_p), ctypes.cast(self.c_row0, ctypes.c_void_p),
                               self.w_stack.data_ptr(), self.w_rows, self.b_stack.data_ptr(), float(first_band(freqs)),
                               int(freqs.numel()))

    def run_forward_rays(self, rays_o, rays_d, z, freqs):
        """nfs_mlp_chain_rays, inference: sampler + encoding + all layers in one launch; z (N,S) -> out (N*S,4)."""
        N, S = z.shape
        out = torch.empty((N * S, 4), device=z.device, dtype=torch.float32)
        if N * S:
            n_hidden = len(self.packed)
            with torch.cuda.device(z.device):
                _lib.call("nfs_mlp_chain_rays", ptr(rays_o), ptr(rays_d), ptr(z), N, S, float(first_band(freqs)),
                          int(freqs.numel()), n_hidden + 1, self.c_k, self.c_n, self.c_act, self.c_row0, ptr(self.w_stack),
                          self.w_rows, ptr(self.b_stack), None, None, None, 0, ptr(out), 4, _stream())
        return out

    def run_forward_fused(self, x16, keep, slot=None, points=None, freqs=None, rays=None):
        """nfs_mlp_chain: all layers in one launch; hidden activations (for wgrad) and their ReLU sign bits
        (for the dgrad chain) are written to HBM only when the backward pass will need them.
        slot = (StepSession, first row): they go into the session's arenas at that row instead of fresh tensors.
        points (+ freqs): training forward with the encoding done in the kernel (nfs_mlp_chain_points_train); x16 is
        then the [P of ceil128(P) rows, k0] buffer that RECEIVES the encoded operand (layer 0's wgrad reads it)."""
        P = x16.shape[0]
        dev = x16.device
        n_hidden = len(self.packed)
        out = torch.empty((P, 4), device=dev, dtype=torch.float32)
        save, bits, rows = None, None, 0
        if slot is not None:
            sess, r0 = slot
            rows = sess.rows_cap                          # per-layer stride of the arenas
            save, bits = sess.save[:, r0:], sess.bits[:, r0:]       # views: data_ptr = row r0 of layer 0
        elif keep:
            rows = _ceil_to(P, 128)
            save = torch.empty((n_hidden, rows, self.h_pad), device=dev, dtype=torch.bfloat16)
            bits = torch.empty((n_hidden, rows, 8), device=dev, dtype=torch.int32)
        if P and rays is not None:
            ro, rd, z = rays
            with torch.cuda.device(dev):
                _lib.call("nfs_mlp_chain_rays", ptr(ro), ptr(rd), ptr(z), z.shape[0], z.shape[1], float(first_band(freqs)),
                          int(freqs.numel()), n_hidden + 1, self.c_k, self.c_n, self.c_act, self.c_row0, ptr(self.w_stack),
                          self.w_rows, ptr(self.b_stack), ptr(x16), ptr(save), ptr(bits), rows, ptr(out), 4, _stream())
        elif P and points is not None:
            with torch.cuda.device(dev):
                _lib.call("nfs_mlp_chain_points_train", ptr(points), float(first_band(freqs)), int(freqs.numel()), P,
                          n_hidden + 1, self.c_k, self.c_n, self.c_act, self.c_row0, ptr(self.w_stack), self.w_rows,
                          ptr(self.b_stack), ptr(x16), ptr(save), ptr(bits), rows, ptr(out), 4, _stream())
        elif P:
            with torch.cuda.device(dev):
                _lib.call("nfs_mlp_chain", ptr(x16), P, n_hidden + 1, self.c_k, self.c_n, self.c_act, self.c_row0,
                          ptr(self.w_stack), self.w_rows, ptr(self.b_stack), None, 0, None, ptr(save), ptr(bits), rows,
                          ptr(out), 4, _stream())
        if slot is not None:
            return out, [x16], None
        acts = [x16] + ([save[i, :P] for i in range(n_hidden)] if keep else [])
        return out, acts, (save, bits) if keep else None

    def run_forward_points(self, pts, freqs):
        """nfs_mlp_chain_points: encoding + all layers in one launch (inference only; nothing saved)."""
        P = pts.shape[0]
        out = torch.empty((P, 4), device=pts.device, dtype=torch.float32)
        if P:
            n_hidden = len(self.packed)
            with torch.cuda.device(pts.device):
                _lib.call("nfs_mlp_chain_points", ptr(pts), float(first_band(freqs)), int(freqs.numel()), P, n_hidden + 1,
                          self.c_k, self.c_n, self.c_act, self.c_row0, ptr(self.w_stack), self.w_rows, ptr(self.b_stack),
                          ptr(out), 4, _stream())
        return out

    def dgrad_chain_fused(self, dy, bits, P):
        """All dgrad GEMMs of the backward pass in one launch (nfs_mlp_chain with act 4 reading the forward
        chain's ReLU sign bits): returns [n_hidden, rows, h_pad] bf16 whose slice j is
        dL/d(pre-activation of layer n_hidden-1-j)."""
        n_hidden = len(self.packed)
        rows = bits.shape[1]
        out = torch.empty((n_hidden, rows, self.h_pad), device=dy.device, dtype=torch.bfloat16)
        with torch.cuda.device(dy.device):
            _lib.call("nfs_mlp_chain", ptr(dy), P, n_hidden, self.cb_k, self.cb_n, self.cb_act, self.cb_row0,
                      ptr(self.wt_stack), self.wt_rows, None, ptr(bits), rows, self.cb_mask, ptr(out), None, rows,
                      None, 0, _stream())
        return out

    # forward over an already-built bf16 operand; returns (out fp32 [P,4], saved activations)
    def run_forward(self, x16, keep):
        """-> (out, acts, save): `save` is the forward chain's [n_hidden, rows, h_pad] activation tensor
        (fused path) or None (layer-by-layer path)."""
        if getattr(self, "fusable", False) and os.environ.get("NFS_MLP_FUSED", "1") != "0":
            return self.run_forward_fused(x16, keep)
        self.refresh_layers()
        acts = [x16]
        h = x16
        for p in self.packed:
            h, _ = ops.linear_bf16(h, p.w16, p.bias, act=1)
            if keep:
                acts.append(h)
        if not keep:
            acts.append(h)
        _, out = ops.linear_bf16(h, self.head.w16, self.head.bias, act=2, out_bf16=False, out_f32_cols=4)
        return out, acts, None

    def run_backward(self, acts, out, g_out, save_fwd=None):
        """Returns the list of parameter gradients in params() order."""
        m = self.module
        dev = out.device
        ps = self.params()
        sizes = [p.numel() for p in ps]
        flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
        views, off = [], 0
        for p, n in zip(ps, sizes):
            views.append(flat[off:off + n].view(p.shape))
            off += n
        n_layers = len(self.packed)
        h_last = acts[-1]
        dy = act_grad(out, g_out, 2, 64)
        # head: rows 0..2 rgb_out, row 3 sigma_out
        tmp_w = torch.zeros(64, self.h_pad, device=dev, dtype=torch.float32)
        tmp_b = torch.zeros(64, device=dev, dtype=torch.float32)
        ops.wgrad_bf16(h_last, dy, tmp_w, 1, self.h_pad, colsum=tmp_b, colsum_of_v=True)
        hd = self.hidden
        views[2 * n_layers + 0].copy_(tmp_w[3:4, :hd])      # sigma_out.weight
        views[2 * n_layers + 1].copy_(tmp_b[3:4])
        views[2 * n_layers + 2].copy_(tmp_w[0:3, :hd])      # rgb_out.weight
        views[2 * n_layers + 3].copy_(tmp_b[0:3])
        if save_fwd is not None and n_layers >= 2 and os.environ.get("NFS_MLP_FUSED_BWD", "1") != "0":
            P = out.shape[0]
            dys = self.dgrad_chain_fused(dy, save_fwd[1], P)
            # the wgrad launches are independent of each other: alternate them between two streams so that one
            # kernel's tail (SMs draining 256 KB of fp32 red.add each) overlaps the next kernel's ramp-up
            main = torch.cuda.current_stream(dev)
            side = _side_stream(dev)
            side.wait_stream(main)
            for n, i in enumerate(range(n_layers - 1, 0, -1)):
                dyi = dys[n_layers - 1 - i, :P]
                with torch.cuda.stream(side if n & 1 else main):
                    ops.wgrad_bf16(acts[i], dyi, views[2 * i], 1, hd, colsum=views[2 * i + 1], colsum_of_v=True,
                                   m_valid=hd, n_valid=hd)
            with torch.cuda.stream(side):
                w0_t = torch.zeros((self.k0, self.h_pad), device=dev, dtype=torch.float32)     # see StepSession.flush
                ops.wgrad_bf16(dys[n_layers - 1, :P], acts[0], w0_t, 1, self.h_pad, colsum=views[1],
                               colsum_of_v=False, m_valid=hd, n_valid=self.in_dim)
                views[0].copy_(w0_t[:self.in_dim, :hd].t())
            main.wait_stream(side)
            return views
        self.refresh_layers()
        dh, _ = ops.linear_bf16(dy, self.head.w16t, None, act=0, relu_mask_src=h_last)
        for i in range(n_layers - 1, 0, -1):
            x_in = acts[i]                                   # input of layer i (= output of layer i-1)
            if hd == self.h_pad:
                ops.wgrad_bf16(x_in, dh, views[2 * i], 1, hd, colsum=views[2 * i + 1], colsum_of_v=True)
            else:
                ops.wgrad_bf16(x_in, dh, views[2 * i], 1, hd, colsum=views[2 * i + 1], colsum_of_v=True,
                               m_valid=hd, n_valid=hd)
            dh, _ = ops.linear_bf16(dh, self.packed[i].w16t, None, act=0, relu_mask_src=x_in)
        # first layer: its K side (the encoding width) is not a multiple of 128 -> M = out features
        ops.wgrad_bf16(dh, acts[0], views[0], self.in_dim, 1, colsum=views[1], colsum_of_v=False,
                       m_valid=hd, n_valid=self.in_dim)
        return views


class StepSession:
    """One optimisation step's worth of MLP calls sharing arenas, so that the weight gradients of ALL calls of the
    step (the coarse and the fine pass of BASELINE config 3) are computed by ONE nfs_wgrad_bf16 launch per layer
    over the concatenated points, accumulated straight into the optimizer's flat gradient buffer:

      begin(rows)  - zero the flat gradient, rewind the arenas (allocated once: static addresses, graph-safe)
      forward      - each NeRFMLP call takes the next 128-row-aligned range: encoded operand, saved activations,
                     ReLU sign bits land in the arenas (run_forward_fused(slot=...))
      backward     - each call runs its head gradient + dgrad chain into the arenas and returns NO parameter
                     gradients (autograd then has nothing to accumulate or copy)
      flush()      - 1 + n_layers wgrad launches over all rows, alternating between two streams.

    Against per-call wgrads this halves the launches (each pays ~15 us of ramp-up and 148 x 256 KB of fp32
    red.add drain), removes the per-parameter autograd accumulation (19 adds) and the gradient flattening
    (~40 copies).  Rows between a call's P and its 128-row boundary hold zero dY in every layer (the chain reads
    rows >= P as zeros through its tensor map; the head's dY rows are zeroed here), so they add nothing."""

    def __init__(self, plan, optimizer):
        self.plan, self.opt = plan, optimizer
        where = {id(p): i for i, p in enumerate(optimizer.params)}
        offs = [0]
        for n in optimizer.sizes:
            offs.append(offs[-1] + n)
        self.views = []
        for p in plan.params():
            if id(p) not in where:
                raise RuntimeError("StepSession: every parameter of the model must belong to the optimizer")
            i = where[id(p)]
            self.views.append(optimizer.grad[offs[i]:offs[i + 1]].view(p.shape))
        self.rows_cap = 0
        self.cursor = 0

    @staticmethod
    def supported(plan, optimizer):
        return (getattr(plan, "fusable", False) and len(plan.packed) >= 2 and hasattr(optimizer, "grad")
                and os.environ.get("NFS_MLP_FUSED", "1") != "0" and os.environ.get("NFS_MLP_FUSED_BWD", "1") != "0"
                and os.environ.get("NFS_MLP_SESSION", "1") != "0"
                and len(optimizer.params) == len(plan.params())          # the session fills the WHOLE flat gradient
                and all(any(p is q for q in optimizer.params) for p in plan.params()))

    def begin(self, rows):
        plan = self.plan
        rows = _ceil_to(rows, 128) + 128 * 8              # room for the 128-row alignment of up to 8 calls
        dev = self.opt.grad.device
        if rows > self.rows_cap:
            n_hidden = len(plan.packed)
            self.rows_cap = rows
            self.x16 = torch.zeros((rows, plan.k0), device=dev, dtype=torch.bfloat16)
            self.save = torch.zeros((n_hidden, rows, plan.h_pad), device=dev, dtype=torch.bfloat16)
            self.bits = torch.zeros((n_hidden, rows, 8), device=dev, dtype=torch.int32)
            self.dy = torch.zeros((rows, 64), device=dev, dtype=torch.bfloat16)
            self.dys = torch.zeros((n_hidden, rows, plan.h_pad), device=dev, dtype=torch.bfloat16)
            self.quad_flags = torch.zeros(2 * (rows // 512 + 2), device=dev, dtype=torch.int32)
        if getattr(self, "tmp", None) is None:
            # lane-contiguous accumulators of the head's and the first layer's weight gradients: one buffer (one fill per
            # step) and one table-driven launch that adds them into the parameters' own layout (_scatter_head)
            hp, k0, hd, L = plan.h_pad, plan.k0, plan.hidden, len(plan.packed)
            self.tmp = torch.zeros(64 * hp + 64 + k0 * hp, device=dev, dtype=torch.float32)
            self.head_w = self.tmp[:64 * hp].view(64, hp)
            self.head_b = self.tmp[64 * hp:64 * hp + 64]
            self.w0_t = self.tmp[64 * hp + 64:].view(k0, hp)
            v, hw, hb = self.views, self.head_w.data_ptr(), self.head_b.data_ptr()
            table = [[hw + 3 * hp * 4, v[2 * L + 0].data_ptr(), 1, hd, hp, 1, hd, 1],      # sigma_out.weight (head row 3)
                     [hb + 3 * 4, v[2 * L + 1].data_ptr(), 1, 1, 1, 1, 1, 1],
                     [hw, v[2 * L + 2].data_ptr(), 3, hd, hp, 1, hd, 1],                   # rgb_out.weight (head rows 0..2)
                     [hb, v[2 * L + 3].data_ptr(), 1, 3, 3, 1, 3, 1],
                     [self.w0_t.data_ptr(), v[0].data_ptr(), hd, plan.in_dim, 1, hp, plan.in_dim, 1]]   # first layer, transposed back
            self.tmp_clean = True               # every entry the weight gradients touch is cleared by the scatter launch
            self.scatter = torch.tensor(table, dtype=torch.int64, device=dev)
            self.scatter_max = max(r[2] * r[3] for r in table)
        self.cursor = 0
        self.pending = 0
        self.n_calls = 0
        self.k1 = []                  # compositing backwards deferred to flush() (merged route): one C call for the whole backward
        self.dy_written = set()
        # One launch for the whole backward pass (nfs_mlp_backward_fused: dgrad chain on producer CTA pairs, weight
        # gradients on consumer CTA pairs fed through L2) when the step is large enough to fill the persistent grid
        # (3.9 vs 4.1-4.3 ms per cfg 3 step); small steps and other widths take the dgrad chain per
        # call + one weight-gradient launch per layer.  NFS_BWD_MERGED=0 / 1 forces either route.
        force = os.environ.get("NFS_BWD_MERGED", "")
        self.merged = plan.h_pad == 256 and (force not in ("", "0") if force != "" else rows >= 65536)
        self.opt.grad.zero_()
        if not self.tmp_clean:                  # only after a step that never reached flush()
            self.tmp.zero_()
        self.tmp_clean = False
        plan._session = self
        return self

    def abort(self):
        """Leave the session without computing gradients (the step raised)."""
        self.plan._session = None
        self.plan._slot = None

    def take(self, P):
        r0 = self.cursor
        pad = _ceil_to(P, 128)
        if r0 + pad > self.rows_cap:
            raise RuntimeError("StepSession: the step evaluates more points than begin() was told")
        self.cursor += pad
        self.pending += 1
        self.n_calls += 1
        return r0

    def defer_composite_backward(self, r0, rgb_sigma, z_vals, rays_d, g_rgb, g_depth, n_rays, n_samples, white):
        """Merged route: the compositing backward of the call at row r0 (it writes the head's dY rows) is not launched
        now but handed to nfs_render_fused_bwd together with the MLP backward at flush().  Returns False when the
        session does not run the merged route (the caller launches nfs_composite_bwd_dy itself)."""
        if not self.merged:
            return False
        self.k1.append(dict(r0=r0, rgb_sigma=rgb_sigma, z_vals=z_vals, rays_d=rays_d, g_rgb=g_rgb, g_depth=g_depth,
                            n_rays=n_rays, n_samples=n_samples, white=white))
        return True

    def backward_call(self, out, g_out, r0):
        """Head gradient + dgrad chain of one call, into the arenas."""
        plan = self.plan
        P = out.shape[0]
        pad = _ceil_to(P, 128)
        if any(k["r0"] == r0 for k in self.k1):          # deferred: nfs_render_fused_bwd writes (and pads) these rows
            self.pending -= 1
            return
        if pad != P:
            self.dy[r0 + P:r0 + pad].zero_()
        if r0 in self.dy_written:
            dy = self.dy[r0:r0 + P]          # the compositing backward has already written this call's rows (ops.composite_loss)
        else:
            dy = act_grad(out, g_out.contiguous(), 2, 64, dst=self.dy[r0:r0 + P])
        self.pending -= 1
        if self.merged:
            return               # flush() runs the dgrad chain of ALL calls together with the weight gradients
        n_hidden = len(plan.packed)
        with torch.cuda.device(dy.device):
            _lib.call("nfs_mlp_chain", ptr(dy), P, n_hidden, plan.cb_k, plan.cb_n, plan.cb_act, plan.cb_row0,
                      ptr(plan.wt_stack), plan.wt_rows, None, ptr(self.bits[:, r0:]), self.rows_cap, plan.cb_mask,
                      ptr(self.dys[:, r0:]), None, self.rows_cap, None, 0, _stream())

    def flush(self):
        """The weight / bias gradients of every call since begin(), into the optimizer's flat gradient."""
        plan, v = self.plan, self.views
        plan._session = None
        if self.pending != 0:
            raise RuntimeError("StepSession: %d MLP call(s) of this step never received a backward pass" % self.pending)
        T = self.cursor
        if T == 0:
            return
        n_layers, hd = len(plan.packed), plan.hidden
        dev = self.opt.grad.device
        if self.merged:
            return self._flush_merged(T)
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.wgrad_bf16(self.save[n_layers - 1, :T], self.dy[:T], self.head_w, 1, plan.h_pad, colsum=self.head_b,
                           colsum_of_v=True)
            self._scatter_head(4)
        for n, i in enumerate(range(n_layers - 1, 0, -1)):
            with torch.cuda.stream(side if n & 1 else main):
                ops.wgrad_bf16(self.save[i - 1, :T], self.dys[n_layers - 1 - i, :T], v[2 * i], 1, hd, colsum=v[2 * i + 1],
                               colsum_of_v=True, m_valid=hd, n_valid=hd)
        with torch.cuda.stream(side):
            # first layer: M must be the 256-wide side (dY), so the kernel's lanes run along the OUTPUT index; into
            # W's own [out,in] layout that is a stride-in_dim scatter of fp32 atomics - accumulate the transpose
            # (lanes contiguous) and add it back transposed
            if os.environ.get("NFS_WGRAD0_DIRECT", "0") != "0":      # developer A/B switch
                ops.wgrad_bf16(self.dys[n_layers - 1, :T], self.x16[:T], v[0], plan.in_dim, 1, colsum=v[1],
                               colsum_of_v=False, m_valid=hd, n_valid=plan.in_dim)
            else:
                ops.wgrad_bf16(self.dys[n_layers - 1, :T], self.x16[:T], self.w0_t, 1, plan.h_pad, colsum=v[1],
                               colsum_of_v=False, m_valid=hd, n_valid=plan.in_dim)
                with torch.cuda.device(dev):
                    _lib.call("nfs_scatter_add_table", ptr(self.scatter[4:]), 1, self.scatter_max, _stream())
            self.tmp_clean = True
        main.wait_stream(side)


    def _flush_merged(self, T):
        """dgrad chain over all T rows + the 1 + n_layers weight gradients, one persistent launch: producer CTA pairs run
        the chain, consumer CTAs pick each 512-row quad of dY up from L2 as soon as it has been stored."""
        plan, v = self.plan, self.views
        n_layers, hd = len(plan.packed), plan.hidden
        jobs = [dict(u=self.save[n_layers - 1, :T], v=self.dy[:T], dw=self.head_w, ld_m=1, ld_n=plan.h_pad,
                     colsum=self.head_b, colsum_of_v=True)]
        waits = [0]
        for i in range(n_layers - 1, 0, -1):
            jobs.append(dict(u=self.save[i - 1, :T], v=self.dys[n_layers - 1 - i, :T], dw=v[2 * i], ld_m=1, ld_n=hd,
                             colsum=v[2 * i + 1], colsum_of_v=True, m_valid=hd, n_valid=hd))
            waits.append(1)
        jobs.append(dict(u=self.dys[n_layers - 1, :T], v=self.x16[:T], dw=self.w0_t, ld_m=1, ld_n=plan.h_pad, colsum=v[1],
                         colsum_of_v=False, m_valid=hd, n_valid=plan.in_dim))
        waits.append(1)
        arr, n, dev, kept = ops.wgrad_job_array(jobs, "nfs_mlp_backward_fused")
        c_waits = (ctypes.c_int32 * n)(*[waits[i] for i in kept])
        k1 = self.k1
        one_call = (len(k1) == self.n_calls and len(k1) > 0 and os.environ.get("NFS_RENDER_FUSED_BWD", "1") != "0"
                    and all(k["rays_d"].data_ptr() == k1[0]["rays_d"].data_ptr() and k["n_rays"] == k1[0]["n_rays"]
                            and k["white"] == k1[0]["white"] for k in k1))
        with torch.cuda.device(dev):
            if one_call:
                # the whole backward of the step behind ONE C call: compositing backward of every pass (writes the head's dY
                # rows) + dgrad chain + all weight gradients (nfs_render_fused_bwd)
                passes = (_lib.RenderPass * len(k1))()
                for ps, k in zip(passes, k1):
                    ps.rgb_sigma, ps.z_vals, ps.g_rgb = k["rgb_sigma"].data_ptr(), k["z_vals"].data_ptr(), k["g_rgb"].data_ptr()
                    ps.g_depth = k["g_depth"].data_ptr() if k["g_depth"] is not None else None
                    ps.n_samples, ps.row0 = k["n_samples"], k["r0"]
                cb = _lib.ChainBackward()
                addr = lambda a: ctypes.cast(a, ctypes.c_void_p).value
                cb.n_layers, cb.k_dims, cb.n_dims, cb.acts, cb.row0 = n_layers, addr(plan.cb_k), addr(plan.cb_n), addr(plan.cb_act), addr(plan.cb_row0)
                cb.wt_stack_bf16, cb.w_rows = plan.wt_stack.data_ptr(), plan.wt_rows
                cb.relu_bits_in, cb.bits_rows_per_layer, cb.mask_idx = self.bits.data_ptr(), self.rows_cap, addr(plan.cb_mask)
                cb.dys_bf16, cb.save_rows_per_layer = self.dys.data_ptr(), self.rows_cap
                cb.jobs, cb.n_jobs, cb.job_waits = addr(arr), n, addr(c_waits)
                cb.quad_flags, cb.producer_pairs = self.quad_flags.data_ptr(), 0
                _lib.call("nfs_render_fused_bwd", ctypes.byref(passes), len(k1), ptr(k1[0]["rays_d"]), k1[0]["n_rays"],
                          int(k1[0]["white"]), ptr(self.dy), int(self.dy.stride(0)), T, ctypes.byref(cb), _stream())
            else:
                for k in k1:                     # (mixed steps: some calls composited elsewhere) pass by pass
                    P = k["n_rays"] * k["n_samples"]
                    pad = _ceil_to(P, 128)
                    if pad != P:
                        self.dy[k["r0"] + P:k["r0"] + pad].zero_()
                    _lib.call("nfs_composite_bwd_dy", ptr(k["rgb_sigma"]), ptr(k["z_vals"]), ptr(k["rays_d"]), ptr(k["g_rgb"]),
                              ptr(k["g_depth"]), None, k["n_rays"], k["n_samples"], int(k["white"]),
                              ptr(self.dy[k["r0"]:k["r0"] + P]), int(self.dy.stride(0)), _stream())
                _lib.call("nfs_mlp_backward_fused", ptr(self.dy), T, n_layers, plan.cb_k, plan.cb_n, plan.cb_act, plan.cb_row0,
                          ptr(plan.wt_stack), plan.wt_rows, ptr(self.bits), self.rows_cap, plan.cb_mask, ptr(self.dys),
                          self.rows_cap, ctypes.byref(arr), n, c_waits, ptr(self.quad_flags), 0, _stream())
        self.k1 = []
        self._scatter_head()

    def _scatter_head(self, entries=5):
        """head_w / head_b / w0_t -> sigma_out, rgb_out and the first layer's gradient views (one launch)."""
        with torch.cuda.device(self.tmp.device):
            _lib.call("nfs_scatter_add_table", ptr(self.scatter), entries, self.scatter_max, _stream())
        self.tmp_clean = entries == 5


class _G1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, keep, x16, extra, *params):
        """extra: None, or dict(slot=(StepSession, row)|None, points=fp32 [P,3]|None, freqs=...) - where the saved
        tensors go and whether the chain kernel encodes the points itself."""
        extra = extra or {}
        slot, points, freqs, rays = extra.get("slot"), extra.get("points"), extra.get("freqs"), extra.get("rays")
        if slot is not None:
            out, _, _ = plan.run_forward_fused(x16, True, slot=slot, points=points, freqs=freqs, rays=rays)
            ctx.plan, ctx.slot, ctx.fused = plan, slot, True
            ctx.save_for_backward(out)
            return out
        ctx.slot = None
        if points is not None or rays is not None:
            out, acts, save = plan.run_forward_fused(x16, True, points=points, freqs=freqs, rays=rays)
        else:
            out, acts, save = plan.run_forward(x16, keep)
        ctx.plan = plan
        ctx.fused = save is not None
        if keep:
            if ctx.fused:
                ctx.save_for_backward(out, x16, save[0], save[1])
            else:
                ctx.save_for_backward(out, *acts)
        return out

    @staticmethod
    def backward(ctx, g_out):
        if ctx.slot is not None:
            (out,) = ctx.saved_tensors
            sess, r0 = ctx.slot
            sess.backward_call(out, g_out, r0)
            return (None, None, None, None) + (None,) * len(ctx.plan.params())
        if ctx.fused:
            out, x16, sv, bits = ctx.saved_tensors
            P = out.shape[0]
            acts = [x16] + [sv[i, :P] for i in range(sv.shape[0])]
            save = (sv, bits)
        else:
            out, *acts = ctx.saved_tensors
            save = None
        grads = ctx.plan.run_backward(acts, out, g_out.contiguous(), save_fwd=save)
        return (None, None, None, None) + tuple(grads)


def g1_forward(plan, x=None, points=None, freqs=None, rays=None):
    """out [P,4] = [sigmoid rgb | raw sigma].  Either `x` (fp32 [P,in_dim], already encoded, the
    reference's calling convention) or `points` (fp32 [P,3]) + `freqs` (encoding fused into the
    operand build, never materialised in fp32), or `rays` = (rays_o (N,3), rays_d (N,3), z_vals (N,S)) + `freqs`
    (sampler AND encoding fused into the chain kernel: out is (N,S,4))."""
    plan.refresh()
    plan._last_slot = None
    if rays is not None:
        return _g1_forward_rays(plan, rays, freqs)
    src = x if x is not None else points
    ops._need_cuda("NeRFMLP", src)
    if src.requires_grad and torch.is_grad_enabled():
        raise RuntimeError("NeRFMLP: gradients w.r.t. the input coordinates are not supported")
    lead = src.shape[:-1]
    flat = src.reshape(-1, src.shape[-1])
    extra = None
    if x is not None:
        if flat.shape[-1] != plan.in_dim:
            raise RuntimeError("NeRFMLP: expected %d input features, got %d" % (plan.in_dim, flat.shape[-1]))
        x16 = encode_operand(flat, None, plan.k0)
    else:
        width = flat.shape[-1] * (2 * int(freqs.numel()) + 1)
        if width != plan.in_dim:
            raise RuntimeError("NeRFMLP: encoding width %d does not match the first layer (%d)" % (width, plan.in_dim))
        params = plan.params()
        keep = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if (not keep and getattr(plan, "fusable", False) and plan.k0 == 64 and flat.shape[-1] == 3
                and 1 <= int(freqs.numel()) <= 10 and bands_are_octaves(freqs)
                and os.environ.get("NFS_MLP_FUSED", "1") != "0" and os.environ.get("NFS_MLP_FUSED_ENC", "1") != "0"):
            # inference: the encoding is produced inside the chain kernel (no operand tensor in HBM)
            return plan.run_forward_points(ops._f32c(flat), freqs).reshape(*lead, 4)
        P = flat.shape[0]
        fused_ok = (getattr(plan, "fusable", False) and keep and P > 0 and os.environ.get("NFS_MLP_FUSED", "1") != "0")
        sess = getattr(plan, "_session", None) if fused_ok else None
        enc_in_kernel = (fused_ok and len(plan.packed) >= 2 and plan.k0 == 64 and flat.shape[-1] == 3
                         and 1 <= int(freqs.numel()) <= 10 and bands_are_octaves(freqs)
                         and os.environ.get("NFS_MLP_FUSED_ENC", "1") != "0")
        plan._last_slot = None
        if sess is not None:
            r0 = sess.take(P)
            extra = {"slot": (sess, r0)}
            plan._last_slot = (sess, r0)      # pipeline.render_rays hands it to the compositing backward (dY written in place)
            dst = sess.x16[r0:r0 + P]
        elif enc_in_kernel:
            dst = torch.empty((_ceil_to(P, 128), plan.k0), device=flat.device, dtype=torch.bfloat16)[:P]
        else:
            dst = None
        if enc_in_kernel:
            # training: the chain kernel encodes the points itself and stores the operand for layer 0's wgrad
            extra = dict(extra or {}, points=ops._f32c(flat), freqs=freqs)
            x16 = dst
        else:
            x16 = encode_operand(flat, freqs, plan.k0, out=dst)
    params = plan.params()
    keep = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    out = _G1Fn.apply(plan, keep, x16, extra, *params)
    return out.reshape(*lead, 4)


def _g1_forward_rays(plan, rays, freqs):
    ro, rd, z = (ops._f32c(t) for t in rays)
    ops._need_cuda("NeRFMLP", ro, rd, z)
    if z.dim() != 2 or ro.shape != (z.shape[0], 3) or rd.shape != ro.shape:
        raise RuntimeError("NeRFMLP.forward_rays: rays_o / rays_d must be (N,3) and z_vals (N,S)")
    N, S = z.shape
    P = N * S
    keep = torch.is_grad_enabled() and any(p.requires_grad for p in plan.params())
    if not plan.can_encode_in_kernel(freqs) or (keep and len(plan.packed) < 2) or P == 0:
        pts = ro[:, None, :] + rd[:, None, :] * z[:, :, None]            # generic route: materialise the positions
        return g1_forward(plan, points=pts.reshape(-1, 3), freqs=freqs).reshape(N, S, 4)
    if not keep:
        return plan.run_forward_rays(ro, rd, z, freqs).reshape(N, S, 4)
    sess = getattr(plan, "_session", None)
    extra = {"rays": (ro, rd, z), "freqs": freqs}
    if sess is not None:
        r0 = sess.take(P)
        extra["slot"] = plan._last_slot = (sess, r0)
        x16 = sess.x16[r0:r0 + P]
    else:
        x16 = torch.empty((_ceil_to(P, 128), plan.k0), device=z.device, dtype=torch.bfloat16)[:P]
    out = _G1Fn.apply(plan, True, x16, extra, *plan.params())
    return out.reshape(N, S, 4)
