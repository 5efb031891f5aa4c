"""Host-side engine of K3 for the feature-conditioned model (G3 in SURVEY.md 3.4):
nerf_mlp.NeRFWithDINO = NeRFDINOFusion -> DensityMLP -> ColorMLP
(/root/reference/src/models/nerf_mlp.py:41-158, dino_feature_model.py:150-197).

The nn.Modules in models/ keep the reference's fp32 parameters; this file owns the launch
sequence.  Forward, P points:
    c    = [enc(x) | f]                         nfs_posenc_bf16, or nfs_g3_operand when f is looked up in a feature map
                                                (projection + lookup + encoding in one kernel) (dino_feature_model.py:182)
    h2   = relu(W2 relu(W1 c))                  one nfs_mlp_chain for both lines: three layers and the
    g    = softmax(Wb relu(Wa h2))              2-way softmax as its fp32 head (:185,188)
    c'   = [enc(x) g0 | f g1]                   nfs_gate_scale_bf16 on the operand c (:191-195)
    h_n  = density layers(Wo relu(W2 relu(W1 c')))   nfs_mlp_chain (3 + n_density layers) (:195-197, nerf_mlp.py:60)
    dens = relu(w_d h_n), feat = W_f h_n        head of that chain; nfs_linear_bf16 (nerf_mlp.py:61-65)
    rgb  = sigmoid(Wc3 relu(Wc2 relu(Wc1 [feat | enc(d)])))   nfs_posenc_bf16 + nfs_linear_bf16 (K = 320) +
                                                nfs_mlp_chain [Wc2, Wc3 as sigmoid head] (:82-84)
Backward: the same graph in reverse - dgrad GEMMs with the ReLU-backward mask fused in their
epilogues (nfs_mlp_chain act 4 / nfs_linear_bf16 relu_mask_src), nfs_wgrad_bf16 for every weight
and bias (the fusion layers W1, W2 receive both of their uses), nfs_gate_bwd_operand for the softmax
gate.  Nothing here computes on the CPU or in eager PyTorch.
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import ptr
from .mlp import PackedLinear, act_grad, bias_terms, encode_operand, pad_hidden, pad_in, _ceil_to
from .ops import _stream


def _i32arr(vals):
    return (ctypes.c_int32 * len(vals))(*vals)


class _Stack:
    """Row-stacked bf16 weights (+ fp32 biases) of a chain of PackedLinear layers: the operand of
    nfs_mlp_chain.  `transposed` stacks the W^T copies (dgrad chain)."""

    def __init__(self, layers, transposed=False):
        self.layers, self.transposed = layers, transposed
        self.key = None

    def alloc(self, dev):
        """Allocate (zeroed) the stacked operands; shapes depend only on the layer widths."""
        shapes = [((p.k_pad, p.n_pad) if self.transposed else (p.n_pad, p.k_pad)) for p in self.layers]
        rows = sum(sh[0] for sh in shapes)
        if getattr(self, "w", None) is None or self.w.device != dev or self.w.shape[0] != rows:
            self.w = torch.zeros(rows, 256, device=dev, dtype=torch.bfloat16)
            self.b = None if self.transposed else torch.zeros(rows, 8, device=dev, dtype=torch.bfloat16)
            self.rows = rows
            row0, r = [], 0
            for sh in shapes:
                row0.append(r)
                r += sh[0]
            self._row0 = row0
            self.c_row0 = _i32arr(row0)
            self.c_k = _i32arr([sh[1] for sh in shapes])
            self.c_n = _i32arr([sh[0] for sh in shapes])

    def table_rows(self):
        rows = []
        for p, r in zip(self.layers, self._row0):
            sub = 0
            for l in p.linears:
                w, b, n, k = l.weight, l.bias, l.out_features, l.in_features
                if self.transposed:
                    rows.append([w.data_ptr(), n, k, k, self.w.data_ptr(), 256, r, sub, 1, 0])
                else:
                    rows.append([w.data_ptr(), n, k, k, self.w.data_ptr(), 256, r + sub, 0, 0, 0])
                    rows.append([b.data_ptr(), n, 0, 0, self.b.data_ptr(), 0, r + sub, 0, 3, 0])
                sub += n
        return rows

    def refresh(self):
        key = tuple(p._key for p in self.layers)
        if key == self.key:
            return self
        mats = [(p.w16t if self.transposed else p.w16) for p in self.layers]
        dev = mats[0].device
        shapes = [((p.k_pad, p.n_pad) if self.transposed else (p.n_pad, p.k_pad)) for p in self.layers]
        rows = sum(sh[0] for sh in shapes)
        w = torch.zeros(rows, 256, device=dev, dtype=torch.bfloat16)
        b = None if self.transposed else torch.zeros(rows, device=dev, dtype=torch.float32)
        r, row0 = 0, []
        for p, m, sh in zip(self.layers, mats, shapes):
            w[r:r + m.shape[0], :m.shape[1]].copy_(m)
            if b is not None:
                b[r:r + m.shape[0]].copy_(p.bias)
            row0.append(r)
            r += sh[0]
        self.w, self.b, self.rows = w, (None if b is None else bias_terms(b)), rows
        self.c_row0 = _i32arr(row0)
        self.c_k = _i32arr([sh[1] for sh in shapes])
        self.c_n = _i32arr([sh[0] for sh in shapes])
        self.key = key
        return self


class _WiderT:
    """The W^T copy of a packed layer as a dgrad-chain layer whose output is zero-padded to `rows` columns (the chain
    kernel saves layers of 128 or 256 columns; the 192-wide d c' of the first fusion layer becomes 256 wide)."""

    def __init__(self, packed, rows):
        self.packed, self.k_pad, self.n_pad, self.linears = packed, rows, packed.n_pad, packed.linears

    w16t = property(lambda self: self.packed.w16t)
    _key = property(lambda self: self.packed._key)


class _SplitColumns(PackedLinear):
    """A Linear whose input is a concatenation of column blocks that sit at padded offsets in the
    bf16 operand ([feat (h_pad) | enc(d) (64)] of the colour MLP, nerf_mlp.py:83)."""

    def __init__(self, linear, blocks, k_pad, n_pad):
        super().__init__([linear], k_pad, n_pad)
        self.blocks = blocks                      # (first input column, width, first operand column)

    def table_rows(self):
        l = self.linears[0]
        n, k_in = l.out_features, l.in_features
        rows = []
        for c_in, width, c_op in self.blocks:
            src = l.weight.data_ptr() + 4 * c_in
            rows.append([src, n, width, k_in, self.w16.data_ptr(), self.k_pad, 0, c_op, 0, 0])
            rows.append([src, n, width, k_in, self.w16t.data_ptr(), self.n_pad, c_op, 0, 1, 0])
        rows.append([l.bias.data_ptr(), n, 0, 0, self.bias.data_ptr(), 0, 0, 0, 2, 0])
        return rows

    def refresh(self):
        l = self.linears[0]
        key = (str(l.weight.device), l.weight.data_ptr(), l.weight._version, l.bias._version, _epoch())
        if key == self._key:
            return self
        dev = l.weight.device
        if not l.weight.is_cuda:
            raise RuntimeError("nfs_b200: model parameters must live on a CUDA device (no CPU fallback)")
        if self.w16 is None or self.w16.device != dev:
            self.w16 = torch.zeros(self.n_pad, self.k_pad, device=dev, dtype=torch.bfloat16)
            self.w16t = torch.zeros(self.k_pad, self.n_pad, device=dev, dtype=torch.bfloat16)
            self.bias = torch.zeros(self.n_pad, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev), torch.no_grad():
            for c_in, width, c_op in self.blocks:
                blk = l.weight.detach()[:, c_in:c_in + width].float().contiguous()
                _lib.call("nfs_pack_linear_bf16", ptr(blk), l.out_features, width, self.n_pad, self.k_pad, 0, c_op,
                          ptr(self.w16), ptr(self.w16t), _stream())
            self.bias[:l.out_features].copy_(l.bias.detach())
        self._key = key
        return self


class _StackedHeads(PackedLinear):
    """[feature_head ; density_head] with the density row at a padded row offset: only its W^T copy
    is used, as the weight of the dgrad GEMM d h_n = [d feat | d dens] . [W_f ; w_d]."""

    def __init__(self, heads, row_offsets, k_pad, n_pad):
        super().__init__(heads, k_pad, n_pad)
        self.row_offsets = row_offsets

    def alloc(self, dev):
        if self.w16t is None or self.w16t.device != dev:
            self.w16t = torch.zeros(self.k_pad, self.n_pad, device=dev, dtype=torch.bfloat16)

    def table_rows(self):
        return [[l.weight.data_ptr(), l.out_features, l.in_features, l.in_features, self.w16t.data_ptr(), self.n_pad,
                 0, r0, 1, 0] for l, r0 in zip(self.linears, self.row_offsets)]

    def refresh(self):
        params = [p for l in self.linears for p in (l.weight, l.bias)]
        dev = params[0].device
        key = (str(dev), _epoch()) + tuple((p.data_ptr(), p._version) for p in params)
        if key == self._key:
            return self
        self.alloc(dev)
        with torch.cuda.device(dev), torch.no_grad():
            for l, r0 in zip(self.linears, self.row_offsets):
                w = l.weight.detach().float().contiguous()
                _lib.call("nfs_pack_linear_bf16", ptr(w), l.out_features, l.in_features, self.n_pad, self.k_pad, r0, 0,
                          None, ptr(self.w16t), _stream())
        self._key = key
        return self


def _epoch():
    from . import mlp
    return mlp._WEIGHT_EPOCH


def _wg(jobs, u, v, dw, ld_m, ld_n, colsum=None, colsum_of_v=True, m_valid=0, n_valid=0):
    if jobs is None:
        ops.wgrad_bf16(u, v, dw, ld_m, ld_n, colsum=colsum, colsum_of_v=colsum_of_v, m_valid=m_valid, n_valid=n_valid)
    else:
        jobs.append(dict(u=u, v=v, dw=dw, ld_m=ld_m, ld_n=ld_n, colsum=colsum, colsum_of_v=colsum_of_v,
                         m_valid=m_valid, n_valid=n_valid))


def wgrad_layer(x16, blocks, dy16, lin, dw, db, jobs=None):
    """dW[out,in] += dY^T X, db[out] += column sums of dY for one Linear (nfs_wgrad_bf16).
    x16 bf16 [P, *] operand; blocks = [(first input column, width, first operand column, padded width)];
    dy16 bf16 [P, n_pad].  The reduction runs over points, so either operand can take the M side:
    a 128/256-wide input block does (its index is the contiguous one of dW), else dY must be 128/256 wide.
    jobs: a list - the launches are appended to it (ops.wgrad_multi runs them together) instead of issued."""
    out_f, in_f = lin.out_features, lin.in_features
    n_pad = dy16.shape[1]
    first = True
    for c_in, width, c_op, w_pad in blocks:
        if w_pad in (128, 256) and n_pad % 64 == 0:
            _wg(jobs, x16[:, c_op:c_op + w_pad], dy16, dw[:, c_in:], 1, in_f, colsum=db if first else None,
                colsum_of_v=True, m_valid=width, n_valid=out_f)
        elif n_pad in (128, 256):
            for c in range(0, w_pad, 256):
                w = min(256, w_pad - c)
                valid = min(width - c, w)
                if valid <= 0:
                    break
                _wg(jobs, dy16, x16[:, c_op + c:c_op + c + w], dw[:, c_in + c:], in_f, 1,
                    colsum=db if first else None, colsum_of_v=False, m_valid=out_f, n_valid=valid)
                first = False
        else:
            raise RuntimeError("nfs_b200: unsupported layer shape for wgrad (%d -> %d)" % (in_f, out_f))
        first = False


class G3Plan:
    def __init__(self, module):
        m = self.module = module
        fu, dm, cm = m.dino_fusion, m.density_mlp, m.color_mlp
        self.W1, self.W2 = fu.fusion[0], fu.fusion[2]
        self.Wa, self.Wb, self.Wo = fu.attention[0], fu.attention[2], fu.output_proj
        self.Wd = [l for l in dm.density_layers if isinstance(l, nn.Linear)]
        self.Wdh, self.Wf = dm.density_head, dm.feature_head
        self.Wc1, self.Wc2, self.Wc3 = cm.color_layers[0], cm.color_layers[2], cm.color_layers[4]
        self.pos_w, self.D, self.dir_w = m.pos_dim, m.dino_dim, m.dir_dim
        H = self.W1.out_features
        if any(l.out_features != H for l in [self.W2, self.Wo, self.Wf] + self.Wd):
            raise RuntimeError("nfs_b200: NeRFWithDINO layers must share one hidden width")
        self.H, self.hp = H, pad_hidden(H)
        self.k0 = pad_in(self.pos_w + self.D)
        self.kd = pad_in(self.dir_w)
        self.cat_k = self.hp + self.kd
        if self.k0 > 256 or self.cat_k > 320 or 3 + len(self.Wd) > 12 or len(self.Wd) < 1:
            raise RuntimeError("nfs_b200: NeRFWithDINO configuration outside the supported shapes "
                               "(encoding + feature width <= 256, hidden <= 256, 1..9 density layers)")
        self.ap, self.c1p, self.c2p = (pad_hidden(self.Wa.out_features), pad_hidden(self.Wc1.out_features),
                                       pad_hidden(self.Wc2.out_features))
        hp = self.hp
        self.p1 = PackedLinear([self.W1], self.k0, hp)
        self.p2 = PackedLinear([self.W2], hp, hp)
        self.pa = PackedLinear([self.Wa], hp, self.ap)
        self.pb = PackedLinear([self.Wb], self.ap, 64)
        self.po = PackedLinear([self.Wo], hp, hp)
        self.pd = [PackedLinear([l], hp, hp) for l in self.Wd]
        self.pdh = PackedLinear([self.Wdh], hp, 64)
        self.pf = PackedLinear([self.Wf], hp, hp)
        self.pheads = _StackedHeads([self.Wf, self.Wdh], [0, hp], hp, hp + 64)
        self.pc1 = _SplitColumns(self.Wc1, [(0, H, 0), (H, self.dir_w, hp)], self.cat_k, self.c1p)
        self.pc2 = PackedLinear([self.Wc2], self.c1p, self.c2p)
        self.pc3 = PackedLinear([self.Wc3], self.c2p, 64)
        self.all_packed = [self.p1, self.p2, self.pa, self.pb, self.po] + self.pd + [self.pdh, self.pf, self.pheads,
                                                                                    self.pc1, self.pc2, self.pc3]
        # chain A: the fusion layers, the first attention layer (its 64 outputs zero-padded to 128 columns) and the gate
        # logits as a 2-way softmax head; its
        # dgrad chain runs from d(gate logits) back to d(pre-activation of W1): [Wb^T, Wa^T, W2^T], each masked by the
        # sign bits chain A saved for the layer's input
        self.chain_a = _Stack([self.p1, self.p2, self.pa, self.pb])
        self.chain_a_bwd = _Stack([self.pb, self.pa, self.p2], transposed=True)
        self.ca_act, self.cab_act, self.cab_mask = _i32arr([1, 1, 1, 5]), _i32arr([4, 4, 4]), _i32arr([2, 1, 0])
        # chain B: second use of the fusion layers, output_proj, the density layers - and the density head as the
        # chain's fp32 output head when the layer count allows
        self.nb = 3 + len(self.Wd)
        body = [self.p1, self.p2, self.po] + self.pd
        self.head_in_chain = self.nb + 1 <= 12
        self.chain_b = _Stack(body + ([self.pdh] if self.head_in_chain else []))
        # dgrad chain of chain B: from d(h_n) down to d(pre-activation of W1's second use); weight of step t =
        # W^T of forward layer nb-1-t, mask = the input of that layer (ReLU output) or none (output_proj is linear)
        # ... and one more step to d c' itself (no mask: c' is the gated input), the operand of the gate's backward
        back = list(reversed(body[1:])) + [_WiderT(self.p1, pad_hidden(self.k0))]
        self.chain_b_bwd = _Stack(back, transposed=True)
        # chain C: the colour MLP behind its first layer (K = 320 does not fit the chain): [Wc2, Wc3 as sigmoid head]
        self.chain_c = _Stack([self.pc2, self.pc3])
        self.cc_act = _i32arr([1, 3])
        acts, midx = [], []
        for t in range(self.nb - 1):
            j_in = self.nb - 2 - t                     # forward layer whose output feeds layer nb-1-t
            acts.append(0 if j_in == 2 else 4)
            midx.append(j_in)
        self.cb_act, self.cb_mask = _i32arr(acts + [0]), _i32arr(midx + [0])
        self.cbf_act = _i32arr([1, 1, 0] + [1] * len(self.Wd) + ([1] if self.head_in_chain else []))

    # parameters in a fixed order; run_backward returns gradients in this order
    def linears(self):
        return [self.W1, self.W2, self.Wa, self.Wb, self.Wo] + self.Wd + [self.Wdh, self.Wf, self.Wc1, self.Wc2, self.Wc3]

    def stacks(self):
        return (self.chain_a, self.chain_a_bwd, self.chain_b, self.chain_b_bwd, self.chain_c)

    def params(self):
        ps = []
        for l in self.linears():
            ps += [l.weight, l.bias]
        return ps

    def refresh(self):
        """Bring every bf16 operand up to date with the fp32 parameters: ONE nfs_pack_table launch (the table lists
        where each parameter lands in the per-layer copies and in the three stacked chain operands)."""
        params = self.params()
        dev = params[0].device
        if not params[0].is_cuda:
            raise RuntimeError("nfs_b200: model parameters must live on a CUDA device (no CPU fallback)")
        key = (str(dev), _epoch()) + tuple((p.data_ptr(), p._version) for p in params)
        if key == getattr(self, "_key", None):
            return
        simple = all(p.dtype == torch.float32 and p.is_contiguous() for p in params)
        if not simple or os.environ.get("NFS_PACK_STACK", "1") == "0":       # layer by layer (any dtype / layout)
            for p in self.all_packed:
                p.refresh()
            for st in self.stacks():
                st.key = None
                st.refresh()
            self._key, self._place = key, None
            return
        place = (str(dev),) + tuple(p.data_ptr() for p in params)
        if place != getattr(self, "_place", None):
            rows = []
            for p in self.all_packed:
                p.alloc(dev)
                rows += p.table_rows()
            for st in self.stacks():
                st.alloc(dev)
                rows += st.table_rows()
            self._table = torch.tensor(rows, dtype=torch.int64, device=dev)
            self._max_elems = max(r[1] * max(r[2], 1) for r in rows)
            self._place = place
        with torch.cuda.device(dev), torch.no_grad():
            _lib.call("nfs_pack_table", ptr(self._table), self._table.shape[0], self._max_elems, _stream())
        for p in self.all_packed:
            p._key = p.current_key() if hasattr(p, "current_key") and type(p).refresh is PackedLinear.refresh else None
        self._key = key

    def _chain(self, x16, stack, n_layers, acts, P, bits_in=None, mask_idx=None, want_bits=False, head_cols=0):
        """nfs_mlp_chain: every non-head layer's output is saved -> [n_saved, rows, widest saved layer]
        (+ the ReLU sign bits [n_saved, rows, 8] of a forward chain when want_bits; + the fp32 [P, head_cols] output of
        the last layer when head_cols > 0)."""
        rows = _ceil_to(P, 128)
        n_saved = n_layers - 1 if head_cols else n_layers
        width = max(stack.c_n[l] for l in range(n_saved))
        save = torch.empty((n_saved, rows, width), device=x16.device, dtype=torch.bfloat16)
        bits = torch.empty((n_saved, rows, 8), device=x16.device, dtype=torch.int32) if want_bits else None
        out = torch.empty((P, head_cols), device=x16.device, dtype=torch.float32) if head_cols else None
        with torch.cuda.device(x16.device):
            _lib.call("nfs_mlp_chain", ptr(x16), P, n_layers, stack.c_k, stack.c_n, acts, stack.c_row0, ptr(stack.w),
                      stack.rows, ptr(stack.b), ptr(bits_in), 0 if bits_in is None else bits_in.shape[1], mask_idx,
                      ptr(save), ptr(bits), rows, ptr(out), head_cols, _stream())
        res = (save,) + ((bits,) if want_bits else ()) + ((out,) if head_cols else ())
        return res if len(res) > 1 else save

    def gate_scale(self, c16, gate):
        """c' = [enc(x) g0 | f g1] from the operand the first pass consumed (nfs_gate_scale_bf16)."""
        out = torch.empty_like(c16)
        with torch.cuda.device(c16.device):
            _lib.call("nfs_gate_scale_bf16", ptr(c16), c16.stride(0), ptr(gate), c16.shape[0], self.pos_w, self.k0,
                      ptr(out), out.stride(0), _stream())
        return out

    def run_forward(self, x, d, f, freqs_pos, freqs_dir, c16=None):
        """c16: the first operand [enc(x) | f | 0] bf16 [P,k0] when the caller has built it already (nfs_g3_operand:
        projection + feature lookup + encoding in one kernel); x / f are not read then."""
        P = d.shape[0]
        hp = self.hp
        if c16 is None:
            c16 = encode_operand(x, freqs_pos, self.k0, extra=f)
        sa, sa_bits, gate = self._chain(c16, self.chain_a, 4, self.ca_act, P, want_bits=True, head_cols=2)
        c2 = self.gate_scale(c16, gate)
        if self.head_in_chain:
            sb, sb_bits, density = self._chain(c2, self.chain_b, self.nb + 1, self.cbf_act, P, want_bits=True, head_cols=1)
        else:
            sb, sb_bits = self._chain(c2, self.chain_b, self.nb, self.cbf_act, P, want_bits=True)
        hn = sb[self.nb - 1, :P]
        if not self.head_in_chain:
            _, density = ops.linear_bf16(hn, self.pdh.w16, self.pdh.bias, act=1, out_bf16=False, out_f32_cols=1)
        cat16 = torch.empty((P, self.cat_k), device=d.device, dtype=torch.bfloat16)
        ops.linear_bf16(hn, self.pf.w16, self.pf.bias, act=0, out=cat16[:, :hp])
        encode_operand(d, freqs_dir, self.kd, out=cat16[:, hp:])
        k1, _ = ops.linear_bf16(cat16, self.pc1.w16, self.pc1.bias, act=1)
        sc, rgb = self._chain(k1, self.chain_c, 2, self.cc_act, P, head_cols=3)
        k2 = sc[0, :P]
        saved = (c16, sa, sa_bits, gate, c2, sb, density, cat16, k1, k2, rgb, sb_bits)
        return rgb, density, saved

    def run_backward(self, saved, g_rgb, g_density):
        c16, sa, sa_bits, gate, c2, sb, density, cat16, k1, k2, rgb, sb_bits = saved
        P = c16.shape[0]
        dev = c16.device
        hp, H, nb = self.hp, self.H, self.nb
        lins = self.linears()
        ps = self.params()
        sizes = [p.numel() for p in ps]
        # every gradient starts on a 16-byte boundary of one zeroed buffer: the weight-gradient kernels drain their
        # accumulators with the TMA bulk reduction (and run on CTA pairs) only into 16-byte aligned destinations - a
        # two-element bias in front (attention[2].bias) otherwise sends every later layer down the per-thread atomic drain
        offs, off = [], 0
        for n in sizes:
            offs.append(off)
            off += (n + 3) // 4 * 4
        flat = torch.zeros(off, device=dev, dtype=torch.float32)
        views = [flat[o:o + n].view(p.shape) for p, n, o in zip(ps, sizes, offs)]
        gw = {id(l): (views[2 * i], views[2 * i + 1]) for i, l in enumerate(lins)}

        # every weight gradient of the step is collected and issued as ONE multi-job launch at the end (a launch
        # per layer costs ~25 us of mostly fixed time at cfg 4's 32 768 points); operands stay alive in `jobs`
        jobs = [] if os.environ.get("NFS_WGRAD_MULTI", "1") != "0" else None

        def wg(lin, x16, dy16, blocks=None):
            dw, db = gw[id(lin)]
            if blocks is None:
                blocks = [(0, lin.in_features, 0, x16.shape[1])]
            wgrad_layer(x16, blocks, dy16, lin, dw, db, jobs=jobs)

        hn = sb[nb - 1, :P]
        # ---- colour MLP (nerf_mlp.py:82-84)
        if g_rgb is not None:
            dy3 = act_grad(rgb, g_rgb, 3, 64)
            wg(self.Wc3, k2, dy3)
            dk2, _ = ops.linear_bf16(dy3, self.pc3.w16t, None, act=0, relu_mask_src=k2)
            wg(self.Wc2, k1, dk2)
            dk1, _ = ops.linear_bf16(dk2, self.pc2.w16t, None, act=0, relu_mask_src=k1)
            wg(self.Wc1, cat16, dk1, blocks=[(0, H, 0, hp), (H, self.dir_w, hp, self.kd)])
        # ---- heads (nerf_mlp.py:61-65): d h_n = [d feat | d dens] . [W_f ; w_d], masked by h_n > 0
        dcat = torch.empty((P, hp + 64), device=dev, dtype=torch.bfloat16)
        if g_rgb is not None:
            ops.linear_bf16(dk1, self.pc1.w16t[:hp], None, act=0, out=dcat[:, :hp])
            wg(self.Wf, hn, dcat[:, :hp])
        else:
            dcat[:, :hp].zero_()
        if g_density is not None:
            act_grad(density, g_density, 1, 64, dst=dcat[:, hp:])
            wg(self.Wdh, hn, dcat[:, hp:])
        else:
            dcat[:, hp:].zero_()
        g_top, _ = ops.linear_bf16(dcat, self.pheads.w16t, None, act=0, relu_mask_src=hn)
        # ---- density layers, output_proj, second use of fusion[2] (dgrad chain)
        dys = self._chain(g_top, self.chain_b_bwd, nb, self.cb_act, P, bits_in=sb_bits, mask_idx=self.cb_mask)

        def grad_pre(j):          # dL/d(pre-activation of chain-B layer j)
            return g_top if j == nb - 1 else dys[nb - 2 - j, :P]

        chain_lins = [self.W1, self.W2, self.Wo] + self.Wd
        for j in range(nb - 1, 0, -1):
            wg(chain_lins[j], sb[j - 1, :P], grad_pre(j))
        k0_blocks = [(0, self.W1.in_features, 0, self.k0)]
        wg(self.W1, c2, grad_pre(0), blocks=k0_blocks)
        # ---- the gate (dino_feature_model.py:188-195)
        dc2 = dys[nb - 1]                              # d c' = d(pre-activation of W1) . W1, last step of the chain
        dlog = torch.empty((P, 64), device=dev, dtype=torch.bfloat16)
        with torch.cuda.device(dev):
            _lib.call("nfs_gate_bwd_operand", ptr(c16), c16.stride(0), ptr(gate), ptr(dc2), dc2.shape[1], P, self.pos_w,
                      self.pos_w + self.D, 64, ptr(dlog), _stream())
        h1, h2, a16 = sa[0, :P], sa[1, :P], sa[2, :P, :self.ap]
        wg(self.Wb, a16, dlog)
        # d a, d h2, d h1 (pre-activations): one dgrad chain [Wb^T, Wa^T, W2^T] masked by chain A's sign bits
        dya = self._chain(dlog, self.chain_a_bwd, 3, self.cab_act, P, bits_in=sa_bits, mask_idx=self.cab_mask)
        wg(self.Wa, h2, dya[0, :P, :self.ap])
        wg(self.W2, h1, dya[1, :P])
        wg(self.W1, c16, dya[2, :P], blocks=k0_blocks)
        if jobs:
            ops.wgrad_multi(jobs)
        return views


class _G3Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, keep, x, d, f, c16, freqs_pos, freqs_dir, *params):
        rgb, density, saved = plan.run_forward(x, d, f, freqs_pos, freqs_dir, c16=c16)
        if keep:
            ctx.plan = plan
            ctx.save_for_backward(*saved)
        ctx.set_materialize_grads(False)
        return rgb, density

    @staticmethod
    def backward(ctx, g_rgb, g_density):
        if g_rgb is None and g_density is None:
            return (None,) * (8 + len(ctx.plan.params()))
        grads = ctx.plan.run_backward(tuple(ctx.saved_tensors), None if g_rgb is None else g_rgb.contiguous(),
                                      None if g_density is None else g_density.contiguous())
        return (None,) * 8 + tuple(grads)


def _g3_apply(plan, x, d, f, c16):
    m = plan.module
    params = plan.params()
    keep = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return _G3Fn.apply(plan, keep, x, d, f, c16, m.pos_encoder.freq_bands, m.dir_encoder.freq_bands, *params)


def g3_forward_from_map(plan, positions, directions, feature_map, pose_inv, focal, H, W):
    """NeRFWithDINO on points whose image features come from ONE source view (train.py:209-231): projection,
    bilinear lookup and positional encoding run as the producer of the first layer's bf16 operand (nfs_g3_operand),
    the (P,C) fp32 features never exist.  feature_map (1,Hp,Wp,C) | (Hp,Wp,C) fp32; pose_inv (4,4) = inverse pose."""
    from .mlp import bands_are_octaves, freqs_on
    m = plan.module
    ops._need_cuda("NeRFWithDINO", positions, directions, feature_map, pose_inv)
    fm = ops._f32c(feature_map if feature_map.dim() == 3 else feature_map[0])
    Hp, Wp, C = fm.shape
    if C != plan.D or plan.D == 0:
        raise RuntimeError("NeRFWithDINO: the feature map has %d channels, the model expects %d" % (C, plan.D))
    plan.refresh()
    x, d = ops._f32c(positions), ops._f32c(directions)
    P = x.shape[0]
    if P == 0:
        return x.new_zeros((0, 3)), x.new_zeros((0, 1))
    bands = m.pos_encoder.freq_bands
    fr = freqs_on(x.device, bands)
    c16 = torch.empty((P, plan.k0), device=x.device, dtype=torch.bfloat16)
    pinv = ops._f32c(pose_inv)               # (torch.inverse returns a column-major tensor: this is a copy, kept alive here)
    with torch.cuda.device(x.device):
        _lib.call("nfs_g3_operand", ptr(x), ptr(pinv), float(focal), int(H), int(W), ptr(fm), Hp, Wp, C,
                  ptr(fr), int(fr.numel()), int(bands_are_octaves(bands)), P, plan.k0, c16.stride(0), ptr(c16), _stream())
    return _g3_apply(plan, None, d, None, c16)


def g3_forward(plan, positions, directions, dino_features):
    """(rgb (P,3), density (P,1)) = NeRFWithDINO.forward (nerf_mlp.py:134-158)."""
    m = plan.module
    ops._need_cuda("NeRFWithDINO", positions, directions, dino_features)
    for t in (positions, directions, dino_features):
        if t is not None and t.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("NeRFWithDINO: gradients w.r.t. positions / directions / features are not supported")
    if positions.dim() != 2 or positions.shape[-1] != 3 or directions.shape != positions.shape:
        raise RuntimeError("NeRFWithDINO: positions and directions must both be (N,3)")
    P = positions.shape[0]
    if plan.D:
        if dino_features is None or dino_features.shape != (P, plan.D):
            raise RuntimeError("NeRFWithDINO: dino_features must be (N,%d)" % plan.D)
        f = ops._f32c(dino_features)
    else:
        if dino_features is not None and dino_features.shape[-1] != 0:
            raise RuntimeError("NeRFWithDINO: model was built with dino_dim=0 but got features %s"
                               % (tuple(dino_features.shape),))
        f = None
    plan.refresh()
    x, d = ops._f32c(positions), ops._f32c(directions)
    if P == 0:
        z = x.new_zeros((0, 3))
        return z, x.new_zeros((0, 1))
    return _g3_apply(plan, x, d, f, None)


# --------------------------------------------------------------------------------- single layers
class _DenseFn(torch.autograd.Function):
    """One nn.Linear (+ReLU / sigmoid) with fp32 tensors at its boundary and the tcgen05 kernels
    inside: used when DensityMLP / ColorMLP / NeRFDINOFusion are called on their own (the reference
    exposes them as modules; inside NeRFWithDINO the whole-model launch plan above runs instead)."""

    @staticmethod
    def forward(ctx, packed, act, x, weight, bias):
        lin = packed.linears[0]
        x16 = encode_operand(x, None, packed.k_pad)
        _, y = ops.linear_bf16(x16, packed.w16, packed.bias, act=act, out_bf16=False, out_f32_cols=lin.out_features)
        ctx.packed, ctx.act = packed, act
        ctx.save_for_backward(x16, y)
        return y

    @staticmethod
    def backward(ctx, g):
        x16, y = ctx.saved_tensors
        packed, lin = ctx.packed, ctx.packed.linears[0]
        dy = act_grad(y, g.contiguous(), ctx.act, packed.n_pad)
        dw = torch.zeros_like(lin.weight)
        db = torch.zeros_like(lin.bias)
        blocks = []
        for c in range(0, packed.k_pad, 256):
            w = min(256, packed.k_pad - c)
            if lin.in_features - c > 0:
                blocks.append((c, min(lin.in_features - c, w), c, w))
        wgrad_layer(x16, blocks, dy, lin, dw, db)
        dx = None
        if ctx.needs_input_grad[2]:
            parts = []
            for c in range(0, packed.k_pad, 256):        # N <= 256 per dgrad launch
                w = min(256, packed.k_pad - c)
                cols = min(lin.in_features - c, w)
                if cols <= 0:
                    break
                _, part = ops.linear_bf16(dy, packed.w16t[c:c + w], None, act=0, out_bf16=False, out_f32_cols=cols)
                parts.append(part)
            dx = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        return None, None, dx, dw, db


_ACT = {"none": 0, "relu": 1, "sigmoid": 3}


def dense(module, linear, x, act="none"):
    """act(linear(x)) for fp32 x (P,K) on CUDA; per-layer packed operands are cached on `module`."""
    ops._need_cuda(type(module).__name__, x)
    if x.dim() != 2 or x.shape[-1] != linear.in_features:
        raise RuntimeError("%s: expected (N,%d) input, got %s" % (type(module).__name__, linear.in_features,
                                                                 tuple(x.shape)))
    cache = module.__dict__.setdefault("_nfs_packed", {})
    packed = cache.get(id(linear))
    if packed is None:
        k_pad, n_pad = pad_in(linear.in_features), _ceil_to(linear.out_features, 64)
        if k_pad > 320 or n_pad > 256:
            raise RuntimeError("nfs_b200: Linear(%d, %d) is outside the supported shapes (in <= 320, out <= 256)"
                               % (linear.in_features, linear.out_features))
        if k_pad not in (128, 256) and n_pad not in (128, 256):
            n_pad = pad_hidden(linear.out_features)      # wgrad needs a 128/256-wide side
        packed = cache[id(linear)] = PackedLinear([linear], k_pad, n_pad)
    packed.refresh()
    if x.shape[0] == 0:
        return x.new_zeros((0, linear.out_features))
    return _DenseFn.apply(packed, _ACT[act], ops._f32c(x), linear.weight, linear.bias)
