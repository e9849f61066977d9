"""LiDAR lossless geometry codec, integer-only inference -- the caller of the hot path
(drop-in for models/convolutional/lossl_coord_int/model.py: same module tree / state-dict keys, same
`compress(xyz) -> bytes` / `decompress(bytes) -> coords` contract, byte-identical bitstreams).

What changed against the reference, all on the device and all bit-exact:
  * the scale pyramid (`get_bin`) is one scan over the Morton-sorted coordinates per level instead of a
    hash build + 8 GEMM launches, and its sizes come back in ONE host read for all 13 levels;
  * every conv / linear is a single fused kernel (gather + int8 GEMM + bias/PReLU/requant [+ residual]);
    `Linear(C -> 8C)` followed by the child mask evaluates only the occupied (node, child) blocks;
  * logits never leave the GPU: the encoder turns them straight into (start, freq) of the coded symbol,
    the decoder into a device-resident CDF table, and the rANS coder itself runs on the GPU -- no
    n x 255 x 2 B table copy, no per-level host synchronisation on the encoder side.
"""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import bitstream, ops
from ..int_sparse_conv.cuda_ops import (
    SharedFxpShift, SparseResBlockIn32W8Out32, SparseConvIn8W8Out8, SparseConvIn8W8Out32,  # noqa: F401
    SparseConvPReLUIn8W8Out8, SparseConvPReLUIn8W8Out32, PReLUIn32Out32, RequantFxpToScaledInt8,
    LinearIn8W8Out8, LinearIn8W8Out32, LinearPReLUIn8W8Out8, LinearPReLUIn8W8Out32, LinearIn8W8, KernelMap)
from ..sparse_tensor import SparseTensor

_DENSE = (PReLUIn32Out32, RequantFxpToScaledInt8, LinearIn8W8)


@dataclass
class Config:
    """models/convolutional/lossl_coord_int/model_config.py:7-17"""
    channels: int = 256
    max_stride_wo_recurrent: int = 2048
    max_stride: int = 8192
    fea_stride: int = 16
    use_more_ch_for_multi_step_pred: bool = False
    skip_top_scales_num: int = 0


class _Trace:
    """FPCC_TRACE=1: synchronising wall-clock marks of the codec phases (diagnostics only)."""

    def __init__(self):
        import os
        self.on = os.environ.get('FPCC_TRACE', '0') == '1'
        self.t = None
        self.rows = []

    def mark(self, name):
        if not self.on:
            return
        import time
        torch.cuda.synchronize()
        now = time.perf_counter()
        if self.t is not None:
            self.rows.append((name, 1e3 * (now - self.t)))
        self.t = now

    def dump(self, title):
        if self.on and self.rows:
            import sys
            sys.stderr.write(f'[trace] {title}: ' + ', '.join(f'{n} {ms:.1f}' for n, ms in self.rows) + '\n')
        self.rows, self.t = [], None


class SparseSequential(nn.Sequential):
    """model.py:524-534; `sel`/`n_out_rows` are forwarded to the LAST linear (occupied-children form)."""

    def forward(self, input: SparseTensor, sel=None, n_out_rows=None, post_requant=None, skip=0, out=None) -> SparseTensor:
        """`post_requant`: fused second stage of the LAST linear (see LinearIn8W8.forward); `skip`: leading modules
        whose work a producer's fused second stage has already done."""
        x = SparseTensor(input.F, input.C, input.stride, input.spatial_range)
        x._caches = input._caches
        last = len(self) - 1
        for i, module in enumerate(self):
            if i < skip:
                continue
            if isinstance(module, _DENSE):
                if i == last and (sel is not None or post_requant is not None):
                    x.F = module(x.F, sel=sel, n_out_rows=n_out_rows, post_requant=post_requant, out=out)
                else:
                    x.F = module(x.F)
            else:
                x = module(x)
        return x


class Level:
    """One level of the scale pyramid: nodes of stride 2^l in Morton order."""

    def __init__(self, C, occ=None, parent=None, slot=None):
        self.C, self.occ, self.parent, self.slot = C, occ, parent, slot
        self._sel = None
        self._bits = None

    @property
    def n(self):
        return self.C.shape[0]

    def sel(self):
        """pairs (parent row, own row) grouped by child slot -> selection for Linear(C->8C) of the parent level"""
        if self._sel is None:
            self._sel = ops.slot_pairs(self.parent, self.slot)
        return self._sel

    def bits_fxp(self):
        """occupancy bits as Q8.23 features [n,8] (model.py:63: cur_bin << SharedFxpShift)"""
        if self._bits is None:
            self._bits = ops.occ_to_bits(self.occ) << SharedFxpShift
        return self._bits

    def symbols(self):
        return self.occ.to(torch.int32) - 1  # model.py:60


def linear_with_bits(requant: RequantFxpToScaledInt8, linear: LinearIn8W8, f: torch.Tensor, occ: torch.Tensor,
                     prelu: Optional[PReLUIn32Out32] = None, aux_requant=None) -> torch.Tensor:
    """`linear(requant([prelu](cat(f, bits << 23))))` (model.py:63-64, 147, 172) without the concat: the bit
    channels requantise to two constants, so their share of the contraction is a bias row chosen by the
    occupancy byte (LinearIn8W8.forward_with_bits) and the GEMM keeps K = C."""
    q0, q1 = requant.bit_levels()
    if CAT_BITS and aux_requant is None and linear.can_cat_bits():
        # the concatenation in memory, without a copy: the requant writes the first C columns of the [rows, C + 16] buffer
        c = linear.in_ch - 8
        buf = torch.empty((f.shape[0], c + 16), dtype=torch.int8, device=f.device)
        requant(f, prelu=prelu, out=buf[:, :c])
        ops.occ_bits_q8(occ, q0, q1, buf[:, c:])
        return linear.forward_cat_bits(buf)
    return linear.forward_with_bits(requant(f, prelu=prelu), occ, q0, q1, aux_requant=aux_requant)


def _with(f, ref: SparseTensor, C=None, stride=None):
    x = SparseTensor(f, ref.C if C is None else C, ref.stride if stride is None else stride, None)
    x._caches = ref._caches
    return x


_K3 = ((3, 3, 3), (1, 1, 1))
import os as _os
# Dual-output fusion (fpcc_epilogue::aux_out): identical bytes either way.  Measured on B200 (96 frames): stand-alone requants
# 35.0 -> 22.8 ms per step, but the producing conv / linear epilogues +7.1 / +4.7 ms: the step does not move (433 vs 435 ms,
# profiles/r02_dual_outputs.txt) and peak memory grows, so it is OFF by default; FPCC_DUAL_OUTPUTS=1 turns it on.
DUAL_OUTPUTS = _os.environ.get('FPCC_DUAL_OUTPUTS', '0') == '1'
# Linear(cat(F, occupancy bits)) as one GEMM over an in-memory concatenation (K = C + 16) instead of the occupancy row-bias
# form (K = C + 1 KB of bias rows per output row read by the epilogue); identical bytes, FPCC_CAT_BITS=0 for the A/B
CAT_BITS = _os.environ.get('FPCC_CAT_BITS', '1') != '0'


def seed_kernel_map(src_caches, dst_caches, coarse_stride, coarse_occ: torch.Tensor, fine: Level):
    """Octree neighbour finding: if the 3x3x3 kernel map of the parent level (stride `coarse_stride`) is cached, derive
    the map of its child level from it (ops.kmap_from_parent: same table as the hash lookup, no hash table of the fine
    level at all) and store it where the conv layers look for it (`_caches.kmaps[(stride, kernel_size, conv_stride)]`).
    The reference rebuilds a hash map per level (lib/int_sparse_conv/cuda_ops.py:113-131)."""
    if fine.parent is None or fine.slot is None or coarse_occ is None:
        return
    entry = src_caches.kmaps.get((tuple(coarse_stride),) + _K3)
    kmap = entry.get('in_out_maps') if entry is not None else None
    if not isinstance(kmap, KernelMap) or kmap.table.shape[1] != coarse_occ.shape[0]:
        return
    tag = (tuple(s // 2 for s in coarse_stride),) + _K3
    if tag not in dst_caches.kmaps:
        table = ops.kmap_from_parent(kmap.table, coarse_occ.contiguous(), fine.parent.contiguous(), fine.slot.contiguous())
        dst_caches.kmaps[tag] = {'in_out_maps': KernelMap(table)}


class OneScalePredictor(nn.Module):
    """model.py:28-92"""

    def __init__(self, channels, if_upsample=True, allow_single_ch=False):
        super().__init__()
        if allow_single_ch:
            self.dec_init = SparseConvIn8W8Out32(1, channels)
        self.dec = SparseResBlockIn32W8Out32(channels)
        self.pred = SparseSequential(
            RequantFxpToScaledInt8(), SparseConvPReLUIn8W8Out8(channels, channels), LinearIn8W8Out32(channels, 255))
        self.pred[-1].padded_output = True  # logits are only read by the CDF kernels, which take a row pitch
        self.if_upsample = if_upsample
        if if_upsample:
            self.upsample = SparseSequential(
                RequantFxpToScaledInt8(), LinearPReLUIn8W8Out32(channels + 8, channels),
                SparseResBlockIn32W8Out32(channels), RequantFxpToScaledInt8(), LinearIn8W8Out32(channels, channels * 8))
        else:
            self.upsample = None

    def trunk(self, cur: SparseTensor):
        if cur.F.shape[1] == 1:
            cur = self.dec_init(cur)
        cur = self.dec(cur)
        return cur, self.pred(cur).F

    def up(self, cur: SparseTensor, occ, child: Level):
        """upsample branch (model.py:62-69): cat(F, bin << 23) -> Requant -> LinearPReLU(C+8 -> C) -> ResBlock ->
        Requant -> Linear(C -> 8C)[child mask]; the concat and the masked 8C output are never materialised."""
        u = self.upsample
        x = _with(linear_with_bits(u[0], u[1], cur.F, occ), cur)
        post = u[3].as_post_stage(None) if u[2].can_fuse_consumer() else None
        if post is not None:  # the ResBlock's only consumer is u[3]: its requant rides in conv2's epilogue
            return u[4](u[2](x, post_requant=post).F, sel=child.sel(), n_out_rows=child.n)
        x = u[2](x)
        return u[4](u[3](x.F), sel=child.sel(), n_out_rows=child.n)


class OneScaleMultiStepPredictor(nn.Module):
    """model.py:95-213"""

    def __init__(self, channels, pred_steps=2, use_more_ch_for_multi_step_pred=True):
        super().__init__()
        self.pred_steps = pred_steps
        k = (2 ** (pred_steps - 2),) * 3
        if pred_steps == 2:
            self.embed = SparseSequential()
            out_ch = channels
            self.dec = SparseSequential(RequantFxpToScaledInt8(), LinearPReLUIn8W8Out32(channels + 8, out_ch),
                                        SparseResBlockIn32W8Out32(out_ch))
        elif use_more_ch_for_multi_step_pred:
            emb = 64 if pred_steps == 3 else 512
            cin = (channels if pred_steps == 3 else round(channels * 1.25)) + emb
            out_ch = round(channels * 1.25) if pred_steps == 3 else channels * 2
            self.embed = SparseSequential(RequantFxpToScaledInt8(), SparseConvPReLUIn8W8Out32(8, emb, k, k))
            self.dec = SparseSequential(RequantFxpToScaledInt8(), LinearPReLUIn8W8Out32(cin, out_ch),
                                        SparseResBlockIn32W8Out32(out_ch)) if cin != out_ch else SparseResBlockIn32W8Out32(out_ch)
        else:
            assert pred_steps >= 3
            self.embed = SparseSequential(
                RequantFxpToScaledInt8(),
                (SparseConvPReLUIn8W8Out32 if channels >= 256 else SparseConvIn8W8Out32)(8, channels, k, k))
            self.dec = SparseSequential(RequantFxpToScaledInt8(), LinearPReLUIn8W8Out32(channels * 2, channels),
                                        SparseResBlockIn32W8Out32(channels))
            out_ch = channels
        self.pred = nn.ModuleList()
        for idx in range(pred_steps):
            if idx == 0:
                self.pred.append(SparseSequential(
                    RequantFxpToScaledInt8(), SparseConvPReLUIn8W8Out8(out_ch, out_ch), LinearIn8W8Out32(out_ch, channels * 8)))
            elif idx != pred_steps - 1:
                self.pred.append(SparseSequential(
                    PReLUIn32Out32(), RequantFxpToScaledInt8(), LinearPReLUIn8W8Out8(channels + 8, channels),
                    SparseConvPReLUIn8W8Out8(channels, channels), LinearIn8W8Out32(channels, channels * 8)))
            else:
                self.pred.append(SparseSequential(
                    RequantFxpToScaledInt8(), SparseConvPReLUIn8W8Out8(channels, channels), LinearIn8W8Out32(channels, 255)))
                self.pred[-1][-1].padded_output = True

    def run(self, cur: SparseTensor, levels: List[Level]):
        """levels[j] = nodes of stride fea_stride / 2^j for j = 0 .. pred_steps-1 (coarse -> fine); all of
        them carry `occ` (encoder: from the pyramid, decoder: already decoded) except the finest, whose
        occupancy is what this block predicts.  Shared by compress and decompress."""
        S = self.pred_steps
        emb_lv = levels[S - 2]  # the level whose occupancy bits are embedded (cur_bins[1] / cur_bins[-1])
        # Dual outputs (fpcc_epilogue::aux_out): an int32 tensor with a Requant among its consumers leaves its producer in
        # both forms -- dec[1]'s Q8.23 rows + the int8 rows of the ResBlock's input_requant; the ResBlock's output + the
        # int8 rows of the Requant that opens pred[0] -- and those stand-alone requant passes disappear.
        res = self.dec[2] if isinstance(self.dec, SparseSequential) and len(self.dec) == 3 and isinstance(self.dec[2], SparseResBlockIn32W8Out32) else None
        dual_in = res is not None and DUAL_OUTPUTS and isinstance(self.dec[1], LinearIn8W8) and self.dec[1].can_emit_aux()
        dual_out = (res is not None and DUAL_OUTPUTS and res.can_fuse_consumer() and isinstance(self.pred[0][0], RequantFxpToScaledInt8)
                    and res.ch % 16 == 0)
        aux_in = res.input_requant.as_post_stage(None) if dual_in else None
        aux_out = self.pred[0][0].as_post_stage(None) if dual_out else None
        cur_q = None  # pred[0][0](cur.F), emitted by the ResBlock

        def res_block(f, f8):
            nonlocal cur_q
            y = res(_with(f, cur), input_q=f8, aux_requant=aux_out)
            if aux_out is not None:
                y, cur_q = y
            return y

        if len(self.embed) == 0:  # pred_steps == 2: the embedding is the occupancy itself
            f = linear_with_bits(self.dec[0], self.dec[1], cur.F, emb_lv.occ, aux_requant=aux_in)
            f, f8 = f if aux_in is not None else (f, None)
            cur = res_block(f, f8)
        else:
            stride = tuple(s >> (S - 2) for s in cur.stride)
            embed_f = self.embed(_with(emb_lv.bits_fxp(), cur, C=emb_lv.C, stride=stride)).F
            dec = self.dec
            if isinstance(dec, SparseSequential) and isinstance(dec[0], RequantFxpToScaledInt8):
                # Requant(cat(F, embed)) == cat(Requant(F), Requant(embed)): one scalar multiplier for all channels, so the
                # concatenation moves int8 rows (4x fewer bytes) and the Q8.23 cat tensor is never built (model.py:199-201)
                c1, c2 = cur.F.shape[1], embed_f.shape[1]
                if c1 % 16 == 0 and c2 % 16 == 0:  # the two requants write the column halves of one buffer: no copy at all
                    q = torch.empty((cur.F.shape[0], c1 + c2), dtype=torch.int8, device=cur.F.device)
                    dec[0](cur.F, out=q[:, :c1])
                    dec[0](embed_f, out=q[:, c1:])
                else:
                    q = torch.cat([dec[0](cur.F), dec[0](embed_f)], 1)
                f = dec[1](q, aux_requant=aux_in)
                f, f8 = f if aux_in is not None else (f, None)
                cur = res_block(f, f8)
            else:
                cur.F = torch.cat([cur.F, embed_f], 1)
                cur = dec(cur)
        x = cur

        def consumer_stage(j):
            """The Q8.23 output of step j's selection linear has ONE consumer, the [PReLU +] Requant that opens step j+1:
            run it inside the linear's epilogue (identical integers, no int32 round trip through HBM) when the linear
            runs on the tensor cores."""
            if j + 1 >= S:
                return None
            nxt = self.pred[j + 1]
            lin = self.pred[j][-1]
            if ops.gemm_engine(lin.in_ch, lin.out_ch // 8) != 'tc' or lin.out_ch % 128 != 0:
                return None
            return nxt[1].as_post_stage(nxt[0]) if j + 1 != S - 1 else nxt[0].as_post_stage(None)

        def cat_buffer(j):
            """When step j's selection linear emits the int8 input of step j+1's Linear(cat(f, bits)) (fused second stage), it
            writes the first C columns of that linear's [rows, C + 16] buffer directly (output pitch): no copy, no row bias."""
            if not CAT_BITS or j + 1 >= S - 1:
                return None
            lin = self.pred[j + 1][2]
            if not (isinstance(lin, LinearIn8W8) and lin.can_cat_bits()):
                return None
            return torch.empty((levels[j + 1].n, lin.in_ch - 8 + 16), dtype=torch.int8, device=cur.F.device)

        fused = None  # second stage already applied to x.F by the producing linear
        catbuf = None  # [n_j, C + 16] buffer whose first C columns the producing linear has filled
        for j, block in enumerate(self.pred):
            post = consumer_stage(j)
            nxt_buf = cat_buffer(j) if post is not None else None
            out_view = nxt_buf[:, : nxt_buf.shape[1] - 16] if nxt_buf is not None else None
            if j == 0:
                src, skip = (cur, 0) if cur_q is None else (_with(cur_q, cur), 1)
                if S > 1:
                    x = block(src, sel=levels[1].sel(), n_out_rows=levels[1].n, post_requant=post, skip=skip, out=out_view)
                else:
                    x = block(src, skip=skip)
                fused, catbuf = post, nxt_buf
                continue
            lv = levels[j]
            f = x.F  # [n_j, C]: features of the occupied children (selection already applied)
            st = tuple(s >> j for s in cur.stride)
            seed_kernel_map(cur._caches, cur._caches, tuple(s >> (j - 1) for s in cur.stride), levels[j - 1].occ, lv)
            if j != S - 1:
                # PReLU -> Requant -> LinearPReLU(C+8 -> C) on cat(f, bits) -> Conv -> Linear(C -> 8C)[child mask]
                if fused is not None:
                    q0, q1 = block[1].bit_levels()
                    if catbuf is not None:  # f IS catbuf[:, :C], written there by the producing linear
                        ops.occ_bits_q8(lv.occ, q0, q1, catbuf[:, catbuf.shape[1] - 16:])
                        f = block[2].forward_cat_bits(catbuf)
                    else:
                        f = block[2].forward_with_bits(f, lv.occ, q0, q1)
                else:
                    f = linear_with_bits(block[1], block[2], f, lv.occ, prelu=block[0])
                x = block[3](_with(f, cur, C=lv.C, stride=st))
                x.F = block[4](x.F, sel=levels[j + 1].sel(), n_out_rows=levels[j + 1].n, post_requant=post, out=out_view)
            else:
                x = block(_with(f, cur, C=lv.C, stride=st), skip=1 if fused is not None else 0)
            fused, catbuf = post, nxt_buf
        return cur, x.F


class Model(nn.Module):
    """model.py:216-521"""

    def __init__(self, cfg: Optional[Config] = None, device='cuda'):
        super().__init__()
        self.cfg = cfg = cfg if cfg is not None else Config()
        self.device = torch.device(device)
        self.max_downsample_times_wo_recurrent = int(np.log2(cfg.max_stride_wo_recurrent))
        self.max_downsample_times = int(np.log2(cfg.max_stride))
        assert cfg.fea_stride >= 2
        self.blocks_dec = nn.ModuleList()
        for idx in range(self.max_downsample_times_wo_recurrent):
            pred_steps = int(np.log2(cfg.fea_stride)) - idx
            if pred_steps < 1:
                self.blocks_dec.append(OneScalePredictor(cfg.channels, True, False))
            elif pred_steps == 1:
                self.blocks_dec.append(OneScalePredictor(cfg.channels, False, False))
            else:
                self.blocks_dec.append(OneScaleMultiStepPredictor(cfg.channels, pred_steps, cfg.use_more_ch_for_multi_step_pred))
        self.block_dec_recurrent = OneScalePredictor(cfg.channels, True, True)
        cdf1 = np.arange(2, 65537).astype(np.uint16)[None]  # model.py:254-257 (the uint16 wrap is the reference's)
        cdf2 = np.arange(1, 129, dtype=np.uint16)[None] * 512
        cdf1[:, -1] = 65535
        cdf2[:, -1] = 65535
        self.register_buffer('fea_side_info_cdf1', torch.from_numpy(cdf1.copy()), persistent=False)
        self.register_buffer('fea_side_info_cdf2', torch.from_numpy(cdf2.copy()), persistent=False)
        self.eval()

    # ---- parameters -------------------------------------------------------------------------
    def load_numpy_state_dict(self, sd):
        """Accepts {reference state-dict key: ndarray}, e.g. from synth.make_lossl_int_state_dict."""
        own = dict(self.named_buffers())
        missing = [k for k in sd if k not in own]
        assert not missing, f'unexpected keys: {missing[:5]}'
        with torch.no_grad():
            for k, v in sd.items():
                t = torch.from_numpy(np.ascontiguousarray(v))
                assert own[k].shape == t.shape and own[k].dtype == t.dtype, (k, own[k].shape, t.shape, own[k].dtype, t.dtype)
                own[k].copy_(t)
        return self

    # ---- helpers ----------------------------------------------------------------------------
    def _num_levels(self):
        return self.max_downsample_times - self.cfg.skip_top_scales_num

    def _block(self, idx):
        blocks = self.blocks_dec[self.cfg.skip_top_scales_num:]
        return self.block_dec_recurrent if idx > len(blocks) else blocks[idx - 1]

    @staticmethod
    def get_init_pc(xyz: torch.Tensor, stride: int = 1) -> SparseTensor:
        return SparseTensor(torch.ones((xyz.shape[0], 1), dtype=torch.int8, device=xyz.device), xyz, (stride,) * 3)

    def build_pyramid(self, xyz: torch.Tensor) -> List[Level]:
        """get_bin for every level (model.py:261-295, 403-405): one scan per level over the sorted coordinates."""
        L = self._num_levels()
        raw = []
        cur = xyz
        for _ in range(L):
            out_c, occ, par, slot, cnt = ops.downsample(cur)
            n = int(cnt.item())  # the next level is launched on the exact prefix
            cur = out_c[:n]
            raw.append((cur, occ[:n], par, slot))
        levels = [Level(xyz, None, raw[0][2], raw[0][3])]
        for l in range(L):
            par, slot = (raw[l + 1][2], raw[l + 1][3]) if l + 1 < L else (None, None)
            levels.append(Level(raw[l][0], raw[l][1], par, slot))
        return levels

    def _frame_rows(self, C: torch.Tensor, n_frames: int) -> torch.Tensor:
        """row range of every frame in a batch-major coordinate list -> int64 [n_frames+1] (device, no sync)"""
        edges = torch.arange(n_frames + 1, dtype=torch.int32, device=C.device)
        return torch.searchsorted(C[:, 0].contiguous(), edges).to(torch.int64)

    # ---- compress ---------------------------------------------------------------------------
    MAX_CDF = 130  # bottom-coordinate alphabet: the stream stores len(cdf)-2 <= 128 (model.py:371)

    MAX_GROUP_FRAMES = 1022  # the coordinate key packs the frame index into 10 bits (include/fastpcc_b200.h: batch < 1023)

    def _check_group_size(self, B: int):
        if B > self.MAX_GROUP_FRAMES:
            raise ValueError(f'{B} frames in one coding group: the coordinate hash keys hold at most {self.MAX_GROUP_FRAMES} frame '
                             f'indices; code the batch in more groups (n_groups >= {-(-B // self.MAX_GROUP_FRAMES)})')

    def _run_groups(self, fn, items: list, n_groups: int) -> list:
        """Runs `fn` on `n_groups` contiguous slices of `items`, each in its own thread and CUDA stream, so that the
        serial range-coder kernels of one group overlap the tensor-core kernels of another.  Order is preserved."""
        n_groups = max(1, min(n_groups, len(items)))
        if n_groups == 1:
            return fn(items)
        import threading
        bounds = [len(items) * g // n_groups for g in range(n_groups + 1)]
        import os
        split = os.environ.get('FPCC_GROUP_SPLIT')  # experiments: relative slice sizes, e.g. "5,4,3" (first = highest priority)
        if split:
            w = [float(v) for v in split.split(',')]
            if len(w) == n_groups:
                acc, tot = 0.0, sum(w)
                bounds = [0]
                for v in w:
                    acc += v
                    bounds.append(int(round(len(items) * acc / tot)))
                bounds[-1] = len(items)
        out: list = [None] * n_groups
        err: list = []
        cur = torch.cuda.current_stream(self.device)

        def work(g):
            try:
                s = self._side_streams[g]
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    out[g] = fn(items[bounds[g]: bounds[g + 1]])
                    s.synchronize()
            except BaseException as e:  # re-raised on the caller's thread
                err.append(e)

        if not hasattr(self, '_side_streams') or len(self._side_streams) < n_groups:
            # descending stream priorities: the groups start together, but the hardware scheduler serves the pending
            # thread blocks of group 0 first, so the groups drift apart and the latency-bound range-coder phase of one
            # group meets the tensor-core phase of another instead of all groups idling in the same phase
            import os
            lo, hi = (0, 0) if os.environ.get('FPCC_STREAM_PRIO', '1') == '0' else (0, -5)
            self._side_streams = [torch.cuda.Stream(self.device, priority=max(hi, lo - g)) for g in range(n_groups)]
        import sys
        from .. import _lib
        old_interval = sys.getswitchinterval()
        sys.setswitchinterval(1e-4)  # the groups interleave thousands of short launches: hand the GIL over quickly
        # leave one SM per concurrently coded stream of the other groups to the serial range-coder kernels
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        import os
        budget = int(os.environ.get('FPCC_SM_BUDGET', '0')) or max(sms // 2, sms - (len(items) - len(items) // n_groups))
        _lib.load().fpcc_set_sm_budget(budget)
        try:
            threads = [threading.Thread(target=work, args=(g,)) for g in range(n_groups)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        finally:
            sys.setswitchinterval(old_interval)
            _lib.load().fpcc_set_sm_budget(0)
        if err:
            raise err[0]
        return [x for part in out for x in part]

    def compress_batch(self, frames: List[torch.Tensor], n_groups: int = 1) -> List[bytes]:
        """Compresses independent frames together (see `_compress_group`); `n_groups` > 1 codes that many slices of
        the batch concurrently on separate CUDA streams."""
        return self._run_groups(self._compress_group, list(frames), n_groups)

    def decompress_batch(self, streams: List[bytes], n_groups: int = 1) -> List[torch.Tensor]:
        return self._run_groups(self._decompress_group, list(streams), n_groups)

    def roundtrip_batch(self, frames: List[torch.Tensor], n_groups: int = 1):
        """compress + decompress of every frame, pipelined per group: each of the `n_groups` slices is encoded to its byte
        strings and decoded from them on its own stream, so the encoder of one slice (tensor-core bound, no serial
        chain) overlaps the level-by-level decoder of another.  Returns (bitstreams, decoded frames), both in order."""
        res = self._run_groups(lambda part: [(d, r) for d, r in zip(*self._roundtrip_group(part))], list(frames), n_groups)
        return [d for d, _ in res], [r for _, r in res]

    def _roundtrip_group(self, frames):
        data = self._compress_group(frames)
        return data, self._decompress_group(data)

    @torch.no_grad()
    def _compress_group(self, frames: List[torch.Tensor]) -> List[bytes]:
        """One pass of every kernel over the concatenated nodes of all frames (the batch index is part of every
        coordinate key), one rANS stream per frame, all streams coded concurrently.  Each returned bitstream is
        byte-identical to `compress(frame)` of the reference."""
        dev, B, L = self.device, len(frames), self._num_levels()
        self._check_group_size(B)
        tr = _Trace()
        tr.mark('start')
        # model.py:396-398 for all frames at once: per-frame minimum, shift, Morton code (x most significant), then
        # ONE pair of stable sorts (Morton, then frame) instead of a sort per frame -- a segmented sort whatever the
        # coordinate width.  Equal points of a frame cannot occur (the input is a voxel set), so the order is unique.
        for xyz in frames:
            assert xyz.dtype == torch.int32 and xyz.dim() == 2 and xyz.shape[1] == 4
        sizes = [int(f.shape[0]) for f in frames]
        frames = [f.to(dev) for f in frames]
        # per-frame minimum before the concatenation: B small reductions (a scatter-min over all points funnels
        # millions of atomics into 3B addresses: 6 ms per 32 frames)
        off_all = torch.stack([f[:, 1:].amin(0) for f in frames]) if min(sizes) > 0 else None
        xyz = torch.cat(frames) if B > 1 else frames[0].clone()
        frame_id = torch.repeat_interleave(torch.arange(B, dtype=torch.int32, device=dev),
                                           torch.tensor(sizes, device=dev), output_size=sum(sizes))
        if off_all is None:  # an empty frame in the batch: the generic segmented minimum
            big = torch.iinfo(torch.int32).max
            off_all = torch.full((B, 3), big, dtype=torch.int32, device=dev)
            off_all.scatter_reduce_(0, frame_id.long()[:, None].expand(-1, 3), xyz[:, 1:], 'amin')
        xyz[:, 1:] -= off_all[frame_id.long()]
        xyz[:, 0] = frame_id
        if B == 1:
            order = torch.argsort(ops.morton_encode(xyz, col0=1, msb_axis=0))
        else:
            o1 = torch.argsort(ops.morton_encode(xyz, col0=1, msb_axis=0))
            o2 = torch.sort(frame_id[o1].to(torch.int16) if B < 32768 else frame_id[o1], stable=True).indices
            order = o1[o2]
        xyz = ops.gather_rows(xyz.contiguous(), order)
        tr.mark('sort')
        levels = self.build_pyramid(xyz)
        tr.mark('pyramid')

        # ---- bottom coordinates and their per-frame histogram CDF (model.py:407-415), vectorised over frames
        V = self.MAX_CDF
        bot = levels[L].C
        frame_of = bot[:, 0].long().repeat_interleave(3)
        bottom_raw = bot[:, 1:].reshape(-1)                                # int32 [3*n_bottom]
        # the stream stores len(cdf) - 2 under a 128-entry table (model.py:371 asserts len(cdf) - 2 <= 128): a coarser
        # bottom level than that cannot be coded.  Index with the clamped value (no device-side assert that would poison
        # the context of the other coding groups) and raise on the host at the next read-back.
        bottom_max = bottom_raw.max() if bottom_raw.numel() else torch.zeros((), dtype=torch.int32, device=dev)
        bottom = bottom_raw.clamp(max=V - 1)
        counts = torch.zeros((B, V), dtype=torch.int64, device=dev)
        counts.view(-1).index_add_(0, frame_of * V + bottom.long(), torch.ones_like(bottom, dtype=torch.int64))
        n_sym = counts.sum(1)                                              # 3 * bottom points of the frame
        ar = torch.arange(V, device=dev)
        n_cdf = ((counts > 0) * (ar + 1)).amax(1).clamp(min=2)            # bincount(minlength=2) length
        scale = torch.div((65536 - n_cdf) << 8, n_sym, rounding_mode='floor')
        pm = ((counts * scale[:, None]) >> 8) + 1
        pm = torch.where(ar[None] < n_cdf[:, None], pm, torch.zeros_like(pm))
        bcdf = pm.cumsum(1)
        bcdf.scatter_(1, (n_cdf - 1)[:, None], torch.full((B, 1), 65535, dtype=torch.int64, device=dev))
        # coder ranges of the bottom coordinates (shared per-frame table; RansEncoder::encode semantics)
        b_lo = torch.where(bottom == 0, torch.zeros_like(frame_of), bcdf[frame_of, (bottom.long() - 1).clamp(min=0)])
        b_hi = torch.where(bottom.long() == n_cdf[frame_of] - 1, torch.full_like(frame_of, 65536), bcdf[frame_of, bottom.long()])
        e_bot = (b_lo | ((b_hi - b_lo - 1) << 16)).to(torch.int32)
        # side info (rans_encode_fea, model.py:367-375): CDF entries minus one under cdf1, then the length under cdf2
        cdf_vals = (bcdf - 1).to(torch.int32).reshape(-1)                  # symbol = cdf[j]-1 for j < n_cdf-1
        e_cdf_all = ops.table_symbol_ranges(self.fea_side_info_cdf1, cdf_vals.clamp(min=0).contiguous()).view(B, V)
        e_len = ops.table_symbol_ranges(self.fea_side_info_cdf2, (n_cdf - 2).to(torch.int32).contiguous())

        tr.mark('bottom')
        # ---- network, coarse -> fine; per coded level the packed (start, freq) of every node's symbol
        cur = SparseTensor(torch.ones((levels[L].n, 1), dtype=torch.int8, device=dev), levels[L].C, (2 ** L,) * 3)
        seg = []
        for idx in range(L, 0, -1):
            blk = self._block(idx)
            lv = levels[idx]
            if isinstance(blk, OneScalePredictor):
                cur, pred = blk.trunk(cur)
                if idx != 1 and blk.if_upsample:
                    f = blk.up(cur, lv.occ, levels[idx - 1])
                    seed_kernel_map(cur._caches, cur._caches, cur.stride, lv.occ, levels[idx - 1])
                    cur = _with(f, cur, C=levels[idx - 1].C, stride=tuple(s // 2 for s in cur.stride))
            else:
                S = blk.pred_steps
                cur, pred = blk.run(cur, [levels[idx + S - 1 - j] for j in range(S)])
            seg.append((lv, ops.cdf_symbol_ranges(pred, lv.symbols())))

        # ---- per-frame streams, entries in DECODE order: cdf length, cdf values, bottom coords, levels coarse ->
        # fine (the encoder pushes the exact reverse: model.py:442-445 and :367-375).  Assembled on the device.
        tr.mark('network')
        lvl_rows = [self._frame_rows(lv.C, B) for lv, _ in seg]           # [B+1] each
        per_frame = 1 + (n_cdf - 1) + n_sym + sum((r[1:] - r[:-1]) for r in lvl_rows)
        start = torch.cumsum(per_frame, 0) - per_frame                     # stream start of every frame
        total_cap = B * V + 3 * bot.shape[0] + sum(lv.n for lv, _ in seg) + B
        ranges = torch.zeros(total_cap, dtype=torch.int32, device=dev)
        ranges[start] = e_len
        j = ar[None].expand(B, V)
        keep = j < (n_cdf - 1)[:, None]
        ranges[(start[:, None] + 1 + j)[keep]] = e_cdf_all[keep]
        bot_rows = self._frame_rows(bot, B) * 3
        pos_in_frame = torch.arange(bottom.shape[0], device=dev) - bot_rows[frame_of]
        ranges[start[frame_of] + n_cdf[frame_of] + pos_in_frame] = e_bot
        base = start + n_cdf + n_sym
        for (lv, rng), rows in zip(seg, lvl_rows):
            fr = lv.C[:, 0].long()
            ranges[base[fr] + torch.arange(lv.n, device=dev) - rows[fr]] = rng
            base = base + (rows[1:] - rows[:-1])
        rng_off = torch.cat([start, (start[-1] + per_frame[-1])[None]]).contiguous()
        longest, bmax = torch.stack([per_frame.max(), bottom_max.to(per_frame.dtype)]).tolist()
        if bmax > V - 1:
            raise ValueError(f'bottom-level coordinate {bmax} exceeds {V - 1}: the per-frame CDF side info holds at most '
                             f'{V} entries (reference model.py:371); use a larger max_stride or fewer skip_top_scales_num')
        cap = (2 * int(longest) + 64 + 3) & ~3  # <= 2 bytes per entry + the 4-byte state header
        tr.mark('assemble')
        out, out_len = ops.rans_encode(ranges, rng_off, cap)
        lens = out_len.tolist()
        tr.mark('rans')
        heads = torch.cat([off_all.long(), (n_sym // 3)[:, None]], 1).tolist()
        assert min(lens) > 0, 'rANS output buffer overflow'
        # only the written tails of the per-stream slots travel to the host (the slots are sized for the worst case)
        packed = torch.cat([out[b, cap - lens[b]:] for b in range(B)])
        host = torch.empty(packed.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(packed, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        out_h = host.numpy()
        result, at = [], 0
        for b in range(B):
            head = b''.join(int(v).to_bytes(2, 'little') for v in heads[b])
            result.append(head + out_h[at: at + lens[b]].tobytes())
            at += lens[b]
        tr.mark('d2h')
        tr.dump(f'compress B={B}')
        return result

    def compress(self, xyz: torch.Tensor) -> bytes:
        return self.compress_batch([xyz])[0]

    def compress_partitions(self, batched_coord: List[torch.Tensor]) -> bytes:
        out = self.compress_batch(list(batched_coord[1:]))  # model.py:455-463 (partitions are independent streams)
        return bitstream.pack_partitions(out)

    # ---- decompress -------------------------------------------------------------------------
    @torch.no_grad()
    def _decompress_group(self, streams: List[bytes]) -> List[torch.Tensor]:
        dev, B, L, V = self.device, len(streams), self._num_levels(), self.MAX_CDF
        self._check_group_size(B)
        heads = np.array([[int.from_bytes(s[2 * i: 2 * i + 2], 'little') for i in range(4)] for s in streams], dtype=np.int64)
        coord_offset = torch.from_numpy(heads[:, :3].astype(np.int32)).to(dev)
        nb = torch.from_numpy(heads[:, 3]).to(dev)                        # bottom points per frame
        lens = np.array([len(s) - 8 for s in streams], dtype=np.int64)
        blob = np.frombuffer(b''.join(s[8:] for s in streams) + b'\0' * ops.RansDecodeStreams.PAD, dtype=np.uint8)
        byte_off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)[:-1]])).to(dev)
        dec = ops.RansDecodeStreams(torch.from_numpy(blob.copy()).to(dev), byte_off,
                                    torch.from_numpy(lens.astype(np.int32)).to(dev), padded=True)

        def offsets(counts):
            return torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(counts.to(torch.int64), 0)]).contiguous()

        # rans_decode_fea(decode_rounded_min=False), model.py:377-393, for all frames at once
        ones = torch.ones(B, dtype=torch.int64, device=dev)
        cdf_len = dec.decode(self.fea_side_info_cdf2, 128, offsets(ones), B, shared=True).long()      # = len(cdf)-2
        n_vals = cdf_len + 1
        n_vals_h = n_vals.tolist()
        vals = dec.decode(self.fea_side_info_cdf1, 65535, offsets(n_vals), int(sum(n_vals_h)), shared=True)
        table = torch.full((B, V), 65535, dtype=torch.int64, device=dev)
        fr = torch.repeat_interleave(torch.arange(B, device=dev), n_vals)
        col = torch.arange(vals.shape[0], device=dev) - offsets(n_vals)[fr]
        table[fr, col] = vals.long() + 1                                   # np.pad(cdf + 1, (0, 1)); cdf[-1] = 65535
        n_sym = nb * 3
        n_sym_h = (heads[:, 3] * 3).tolist()
        bottom = dec.decode(ops.as_u16(table).contiguous(), V, offsets(n_sym), int(sum(n_sym_h)), rows_per_stream=True,
                            s_per_stream=(cdf_len + 2).to(torch.int32).contiguous())
        fr = torch.repeat_interleave(torch.arange(B, device=dev, dtype=torch.int32), nb)
        C = torch.cat([fr[:, None], bottom.reshape(-1, 3)], 1).contiguous()
        cur = self.get_init_pc(C, 2 ** L)

        def decode_level(logits, lv: Level):
            rows = self._frame_rows(lv.C, B)
            sym = dec.decode(ops.quantize_cdf(logits), logits.shape[1], rows, lv.n)
            lv.occ = (sym + 1).to(torch.uint8)  # occ byte = oct symbol + 1 (model.py:81)
            cc, par, slot, _ = ops.upsample(lv.C, lv.occ)
            return Level(cc, None, par, slot)

        ms_levels: List[Level] = []  # decoded levels from fea_stride downwards, for the multi-step blocks
        lv = Level(C)
        for idx in range(L, 0, -1):
            blk = self._block(idx)
            if isinstance(blk, OneScalePredictor):
                cur, pred = blk.trunk(cur)
                child = decode_level(pred, lv)
                if idx != 1 and blk.if_upsample:
                    f = blk.up(cur, lv.occ, child)
                    nxt = SparseTensor(f, child.C, tuple(s // 2 for s in cur.stride))  # fresh caches (model.py:88-91)
                    seed_kernel_map(cur._caches, nxt._caches, cur.stride, lv.occ, child)
                    cur = nxt
                else:
                    ms_levels = [lv]
                lv = child
            else:
                S = blk.pred_steps
                if len(ms_levels) < S:
                    ms_levels.append(lv)
                for j in range(1, S):  # register the decoded coordinate sets (model.py:190)
                    cur._caches.cmaps.setdefault(tuple(s >> j for s in cur.stride), (ms_levels[j].C, None))
                cur, pred = blk.run(cur, ms_levels[:S])
                lv = decode_level(pred, ms_levels[S - 1])
        assert not dec.error(), 'corrupt bitstream'
        rows = self._frame_rows(lv.C, B).tolist()
        return [lv.C[rows[b]: rows[b + 1], 1:] + coord_offset[b][None] for b in range(B)]

    def decompress(self, compressed_bytes: bytes) -> torch.Tensor:
        return self.decompress_batch([compressed_bytes])[0]

    def decompress_partitions(self, concat_bytes: bytes) -> torch.Tensor:
        parts = bitstream.split_partitions(concat_bytes)  # model.py:510-521; raises on a truncated container
        return torch.cat(self.decompress_batch(parts), 0)
