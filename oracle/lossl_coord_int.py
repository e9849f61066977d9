"""numpy restatement of the reference's integer-only lossless LiDAR geometry codec
(models/convolutional/lossl_coord_int/model.py) on top of oracle/int_ops.py and oracle/rans.py.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity status: PINNED at the codec level -- the bitstreams and
decoded point order of this file equal those of the reference's own Python (unmodified model.py + cuda_ops.py +
compiled range coder, run on the CPU by tests/golden/make_int_codec_golden.py; fixture
tests/golden/int_codec_golden.json, test tests/test_oracle_int_codec_golden.py).  The 20 primitive entry points of
the CUDA extension underneath (oracle/int_ops.py) are pinned to their sources only: the extension cannot be built
offline.  Parameters come in as a plain {state_dict_key: ndarray} mapping with the reference's
own key names (lib/int_sparse_conv/cuda_ops.py:194-206, 476-481, 516-528).
"""
import numpy as np

from . import int_ops as K
from .rans import RansDecoder, RansEncoder


class SparseTensor:
    """Minimal stand-in for torchsparse.SparseTensor as used by the reference
    (F, C, stride, shared caches: cuda_ops.py:324-361)."""

    def __init__(self, F, C, stride, caches=None):
        self.F, self.C = F, C
        self.stride = tuple(stride) if isinstance(stride, (tuple, list)) else (int(stride),) * 3
        self.caches = caches if caches is not None else {'cmaps': {}, 'kmaps': {}}


def morton_xmajor(xyz):
    """morton_encode_magicbits(xyz, inverse=True) (lib/space_filling_curves/__init__.py:65-88):
    x is the most significant interleaved bit, z the least."""
    def split3(a):
        x = a.astype(np.uint64)
        x = (x | (x << np.uint64(32))) & np.uint64(0x1f00000000ffff)
        x = (x | (x << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
        x = (x | (x << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
        x = (x | (x << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
        x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
        return x
    return (split3(xyz[:, 2]) | (split3(xyz[:, 1]) << np.uint64(1)) | (split3(xyz[:, 0]) << np.uint64(2))).astype(np.int64)


# ---------------------------------------------------------------------------------------------
# layers (lib/int_sparse_conv/cuda_ops.py)
# ---------------------------------------------------------------------------------------------

class _P:
    def __init__(self, sd, prefix):
        self.sd, self.prefix = sd, prefix

    def get(self, name):
        return self.sd[self.prefix + name]

    def has(self, name):
        return (self.prefix + name) in self.sd

    def sub(self, name):
        return _P(self.sd, self.prefix + name + '.')


class Requant:  # RequantFxpToScaledInt8, cuda_ops.py:473-509
    def __init__(self, p):
        self.mul, self.shift, self.zp = p.get('requant_mul'), int(p.get('requant_shift')[0]), p.get('int_zero_point_out')

    def __call__(self, f):
        return K.requant(f, np.repeat(self.mul, f.shape[1]), self.zp, K.SharedFxpShift + self.shift, np.int8)


class PReLU32:  # PReLUIn32Out32, cuda_ops.py:458-470
    def __init__(self, p):
        self.slope = p.get('slope')

    def __call__(self, f):
        return K.prelu(f, self.slope)


class Linear:  # LinearIn8W8 family, cuda_ops.py:512-635
    def __init__(self, p, out_scaled_int):
        self.w, self.b = p.get('weight'), p.get('bias')
        self.slope = p.get('slope') if p.has('slope') else None
        self.mul, self.shift, self.zp = p.get('requant_mul'), int(p.get('requant_shift')[0]), p.get('int_zero_point_out')
        self.out8 = out_scaled_int

    def __call__(self, f):
        mm = K.gemm_int8(f, self.w, self.b)
        if self.out8:
            return K.requant(mm, self.mul, self.zp, self.shift, np.int8, slope=self.slope)
        return K.requant(mm, self.mul, self.zp, self.shift - K.SharedFxpShift, np.int32, slope=self.slope)


class Conv:  # SparseConvIn8Out8 family, cuda_ops.py:189-405
    def __init__(self, p, kernel_size, stride, out_scaled_int):
        self.w, self.b = p.get('weight'), p.get('bias')
        self.slope = p.get('slope') if p.has('slope') else None
        self.mul, self.shift, self.zp = p.get('requant_mul'), int(p.get('requant_shift')[0]), p.get('int_zero_point_out')
        self.zp_comp = p.get('int_zero_point_in_comp') if p.has('int_zero_point_in_comp') else None
        self.ks, self.st, self.out8 = tuple(kernel_size), tuple(stride), out_scaled_int

    def __call__(self, x: SparseTensor) -> SparseTensor:
        caches = x.caches
        tag = (x.stride, self.ks, self.st)
        maps = caches['kmaps'].get(tag)
        if self.st == (1, 1, 1):
            out_stride, out_c, same = x.stride, x.C, True
        else:
            same = False
            out_stride = tuple(a * b for a, b in zip(x.stride, self.st))
            if out_stride in caches['cmaps']:
                out_c = caches['cmaps'][out_stride]
            else:  # cuda_ops.py:341-344 (torch.unique(dim=0) sorts rows lexicographically)
                sh = self.st[0].bit_length() - 1
                oc = x.C.copy()
                oc[:, 1:] >>= sh
                out_c = np.unique(oc, axis=0)
        acc, maps = K.sparse_conv_in8w8out32(x.F, self.w, x.C, out_c, self.ks, self.st, maps, self.zp_comp, same)
        caches['kmaps'].setdefault(tag, maps)
        caches['cmaps'].setdefault(x.stride, x.C)
        caches['cmaps'].setdefault(out_stride, out_c)
        if self.out8:
            f = K.requant(acc, self.mul, self.zp, self.shift, np.int8, bias=self.b, slope=self.slope)
        else:
            f = K.requant(acc, self.mul, self.zp, self.shift - K.SharedFxpShift, np.int32, bias=self.b, slope=self.slope)
        return SparseTensor(f, out_c, out_stride, caches)


class ResBlock:  # SparseResBlockIn32W8Out32, cuda_ops.py:62-92
    def __init__(self, p):
        self.rq = Requant(p.sub('input_requant'))
        self.c1 = Conv(p.sub('conv_prelu'), (3, 3, 3), (1, 1, 1), True)
        self.c2 = Conv(p.sub('conv2'), (3, 3, 3), (1, 1, 1), False)
        self.act = PReLU32(p.sub('prelu'))

    def __call__(self, x):
        y = self.c2(self.c1(SparseTensor(self.rq(x.F), x.C, x.stride, x.caches)))
        s = K._wrap32(x.F.astype(np.int64) + y.F.astype(np.int64))  # torch int32 add wraps
        return SparseTensor(self.act(s), x.C, x.stride, x.caches)


class Seq:  # SparseSequential, lossl_coord_int/model.py:524-534
    def __init__(self, mods):
        self.mods = mods

    def __call__(self, x):
        x = SparseTensor(x.F, x.C, x.stride, x.caches)
        for m in self.mods:
            if isinstance(m, (Requant, PReLU32, Linear)):
                x.F = m(x.F)
            else:
                x = m(x)
        return x


BIN2OCT = np.arange(7, -1, -1, dtype=np.int64)  # model.py:243
UNFOLD = np.array([(0, (k >> 2) & 1, (k >> 1) & 1, k & 1) for k in range(8)], dtype=np.int32)[None]  # model.py:244-246


def _oct_of(bits):
    return ((bits.astype(np.int64) << BIN2OCT[None]).sum(1) - 1).astype(np.uint16)  # model.py:60


def _bits_of(oct_):
    return (((oct_.astype(np.int64)[:, None] + 1) >> BIN2OCT[None]) & 1).astype(bool)  # model.py:81


class OneScalePredictor:  # model.py:28-92
    def __init__(self, p, ch, if_upsample, allow_single_ch):
        self.dec_init = Conv(p.sub('dec_init'), (3, 3, 3), (1, 1, 1), False) if allow_single_ch else None
        self.dec = ResBlock(p.sub('dec'))
        q = p.sub('pred')
        self.pred = Seq([Requant(q.sub('0')), Conv(q.sub('1'), (3, 3, 3), (1, 1, 1), True), Linear(q.sub('2'), False)])
        self.if_upsample = if_upsample
        if if_upsample:
            u = p.sub('upsample')
            self.upsample = Seq([Requant(u.sub('0')), Linear(u.sub('1'), False), ResBlock(u.sub('2')),
                                 Requant(u.sub('3')), Linear(u.sub('4'), False)])

    def _trunk(self, cur):
        if cur.F.shape[1] == 1:
            cur = self.dec_init(cur)
        cur = self.dec(cur)
        return cur, self.pred(cur).F

    def _up(self, cur, bits):
        cur.F = np.concatenate([cur.F, bits.astype(np.int32) << K.SharedFxpShift], 1)
        cur = self.upsample(cur)
        return cur, cur.F.reshape(cur.F.shape[0], 8, cur.F.shape[1] // 8)[bits.astype(bool)]

    def compress(self, cur, up_ref, cur_bin, if_upsample):
        cur, pred = self._trunk(cur)
        oct_ = _oct_of(cur_bin)
        if if_upsample:
            cur, f = self._up(cur, cur_bin)
            cur = SparseTensor(f, up_ref.C, tuple(s // 2 for s in cur.stride), up_ref.caches)
        return cur, pred, oct_

    def decompress(self, cur, decode_oct, if_upsample):
        cur, pred = self._trunk(cur)
        bits = _bits_of(decode_oct(pred))
        if if_upsample:
            cur, f = self._up(cur, bits)
            new_c = cur.C[:, None].copy()
            new_c[..., 1:] <<= 1
            cur = SparseTensor(f, (new_c + UNFOLD)[bits], tuple(s // 2 for s in cur.stride))  # fresh caches (model.py:88-91)
        return cur, bits


class OneScaleMultiStepPredictor:  # model.py:95-213
    def __init__(self, p, ch, steps, use_more):
        self.steps = steps
        e, d = p.sub('embed'), p.sub('dec')
        if steps == 2:
            self.embed = Seq([])
            out_ch = ch
            self.dec = Seq([Requant(d.sub('0')), Linear(d.sub('1'), False), ResBlock(d.sub('2'))])
        else:
            k = 2 ** (steps - 2)
            self.embed = Seq([Requant(e.sub('0')), Conv(e.sub('1'), (k,) * 3, (k,) * 3, False)])
            if use_more:
                emb = 64 if steps == 3 else 512
                cin = (ch if steps == 3 else round(ch * 1.25)) + emb
                out_ch = round(ch * 1.25) if steps == 3 else ch * 2
                self.dec = Seq([Requant(d.sub('0')), Linear(d.sub('1'), False), ResBlock(d.sub('2'))]) \
                    if cin != out_ch else ResBlock(d)
            else:
                out_ch = ch
                self.dec = Seq([Requant(d.sub('0')), Linear(d.sub('1'), False), ResBlock(d.sub('2'))])
        self.pred = []
        for i in range(steps):
            q = p.sub(f'pred.{i}')
            if i == 0 or i == steps - 1:
                self.pred.append(Seq([Requant(q.sub('0')), Conv(q.sub('1'), (3, 3, 3), (1, 1, 1), True),
                                      Linear(q.sub('2'), False)]))
            else:
                self.pred.append(Seq([PReLU32(q.sub('0')), Requant(q.sub('1')), Linear(q.sub('2'), True),
                                      Conv(q.sub('3'), (3, 3, 3), (1, 1, 1), True), Linear(q.sub('4'), False)]))

    def compress(self, cur, cur_bins):
        ein = SparseTensor(cur_bins[1].F.astype(np.int32) << K.SharedFxpShift, cur_bins[1].C, cur_bins[1].stride, cur.caches)
        cur.F = np.concatenate([cur.F, self.embed(ein).F], 1)
        cur = self.dec(cur)
        pred = self.pred[0](cur)
        for i in range(1, self.steps):
            m = cur_bins[-i].F.astype(bool)
            pred.F = pred.F.reshape(pred.F.shape[0], 8, pred.F.shape[1] // 8)[m]
            if i != self.steps - 1:
                pred.F = np.concatenate([pred.F, cur_bins[-i - 1].F.astype(np.int32) << K.SharedFxpShift], 1)
            pred.C, pred.stride = cur_bins[-i - 1].C, cur_bins[-i - 1].stride
            pred = self.pred[i](pred)
        return cur, pred.F, _oct_of(cur_bins[0].F)

    def decompress(self, cur, cur_bins, top_rec, top_stride, decode_oct):
        if len(cur_bins) == 1:
            top_rec, top_stride = cur.C, cur.stride[0]
        t = top_rec[:, None].copy()
        t[..., 1:] <<= 1
        top_rec = (t + UNFOLD)[cur_bins[-1]]
        top_stride //= 2
        cur.caches['cmaps'][(top_stride,) * 3] = top_rec
        ein = SparseTensor(cur_bins[-1].astype(np.int32) << K.SharedFxpShift,
                           cur.caches['cmaps'][(top_stride * 2,) * 3], (top_stride * 2,) * 3, cur.caches)
        cur.F = np.concatenate([cur.F, self.embed(ein).F], 1)
        cur = self.dec(cur)
        pred = self.pred[0](cur)
        for i in range(1, self.steps):
            pred.F = pred.F.reshape(pred.F.shape[0], 8, pred.F.shape[1] // 8)[cur_bins[i - 1]]
            if i != self.steps - 1:
                pred.F = np.concatenate([pred.F, cur_bins[i].astype(np.int32) << K.SharedFxpShift], 1)
            pred.stride = tuple(s // 2 for s in pred.stride)
            pred.C = cur.caches['cmaps'][pred.stride]
            pred = self.pred[i](pred)
        return cur, _bits_of(decode_oct(pred.F)), top_rec, top_stride


# ---------------------------------------------------------------------------------------------
# Model (model.py:216-521)
# ---------------------------------------------------------------------------------------------

class Model:
    def __init__(self, state_dict, channels=256, max_stride_wo_recurrent=2048, max_stride=8192, fea_stride=16,
                 use_more_ch_for_multi_step_pred=False, skip_top_scales_num=0):
        self.skip = skip_top_scales_num
        self.n_wo = int(np.log2(max_stride_wo_recurrent))
        self.n_ds = int(np.log2(max_stride))
        root = _P(state_dict, '')
        self.blocks = []
        for i in range(self.n_wo):
            steps = int(np.log2(fea_stride)) - i
            p = root.sub(f'blocks_dec.{i}')
            if steps < 1:
                self.blocks.append(OneScalePredictor(p, channels, True, False))
            elif steps == 1:
                self.blocks.append(OneScalePredictor(p, channels, False, False))
            else:
                self.blocks.append(OneScaleMultiStepPredictor(p, channels, steps, use_more_ch_for_multi_step_pred))
        self.recurrent = OneScalePredictor(root.sub('block_dec_recurrent'), channels, True, True)
        cdf1 = np.arange(2, 65537).astype(np.uint16)[None]  # model.py:254-257 (wraps, as in the reference)
        cdf2 = (np.arange(1, 129, dtype=np.uint16)[None] * 512)
        cdf1[:, -1] = 65535
        cdf2[:, -1] = 65535
        self.cdf1, self.cdf2 = cdf1, cdf2
        self.trace = None  # filled by compress(): per-level (cdf, symbols) for finer-grained tests

    # -- model.py:261-295; occupancy bits via the eye-shaped fold2bin conv
    def get_bin(self, x: SparseTensor) -> SparseTensor:
        oc = x.C.copy()
        oc[:, 1:] >>= 1
        keep = np.ones(oc.shape[0], dtype=bool)
        keep[1:] = (oc[1:] != oc[:-1]).any(1)  # torch.unique_consecutive(dim=0)
        oc = oc[keep]
        ones = np.ones((x.C.shape[0], 1), dtype=np.int8)
        w = np.eye(8, dtype=np.int8).reshape(8, 8, 1)
        f, maps = K.sparse_conv_in8w8out32(ones, w, x.C, oc, (2, 2, 2), (2, 2, 2), None, None, True)
        out_stride = tuple(s * 2 for s in x.stride)
        if x.stride != (1, 1, 1):
            x.caches['kmaps'].setdefault((x.stride, (2, 2, 2), (2, 2, 2)), maps)
            x.caches['cmaps'].setdefault(x.stride, x.C)
        x.caches['cmaps'].setdefault(out_stride, oc)
        return SparseTensor(f, oc, out_stride, x.caches)

    def _levels(self):
        L = self.n_ds - self.skip
        blocks = self.blocks[self.skip:]
        for idx in range(L, 0, -1):
            yield idx, (self.recurrent if idx > len(blocks) else blocks[idx - 1])

    def compress(self, xyz: np.ndarray) -> bytes:
        assert xyz.dtype == np.int32 and xyz.shape[1] == 4
        off = xyz[:, 1:].min(0)
        xyz = xyz - np.concatenate([[0], off]).astype(np.int32)[None]
        xyz = xyz[np.argsort(morton_xmajor(xyz[:, 1:]), kind='stable')]
        org = SparseTensor(np.ones((xyz.shape[0], 1), dtype=np.int8), xyz, 1)
        L = self.n_ds - self.skip
        sl = [org]
        for _ in range(L):
            sl.append(self.get_bin(sl[-1]))
        bottom = sl[-1].C[:, 1:].reshape(-1)
        counts = np.bincount(bottom, minlength=2).astype(np.int64)
        pm = ((counts * (((65536 - counts.shape[0]) << 8) // bottom.size)) >> 8) + 1  # model.py:409-415
        bcdf = np.cumsum(pm)
        bcdf[-1] = 65535
        bcdf = bcdf.astype(np.uint16)
        cur = SparseTensor(np.ones((sl[-1].C.shape[0], 1), dtype=np.int8), sl[-1].C, 2 ** L, org.caches)
        cached = []
        for idx, blk in self._levels():
            if isinstance(blk, OneScalePredictor):
                cur, pred, oct_ = blk.compress(cur, sl[idx - 1], sl[idx].F, idx != 1 and blk.if_upsample)
            else:
                cur, pred, oct_ = blk.compress(cur, sl[idx: idx + blk.steps])
            cached.append((K.batch_quantize_pmf(pred), oct_))
        self.trace = {'levels': cached, 'bottom': (bcdf, bottom.astype(np.uint16)), 'pyramid': sl}
        enc = RansEncoder(32 * 1024 * 1024)
        for cdf, oct_ in reversed(cached):  # model.py:442-444
            enc.encode(cdf, oct_)
        # rans_encode_fea, model.py:367-375
        enc.encode(bcdf[None], bottom.astype(np.uint16))
        enc.encode(self.cdf1, (bcdf[:-1] - 1).astype(np.uint16))
        assert len(bcdf) - 2 <= self.cdf2.shape[1]
        enc.encode(self.cdf2, np.array([len(bcdf) - 2], dtype=np.uint16))
        head = b''.join(int(v).to_bytes(2, 'little') for v in off.tolist())
        head += int(bottom.shape[0] // 3).to_bytes(2, 'little')
        return head + enc.flush()

    def decompress(self, data: bytes) -> np.ndarray:
        off = np.array([int.from_bytes(data[2 * i: 2 * i + 2], 'little') for i in range(3)], dtype=np.int32)
        nb = int.from_bytes(data[6:8], 'little')
        dec = RansDecoder()
        payload = data[8:]  # the reference's decoder keeps a pointer into these bytes: hold them until decoding is done
        dec.flush(payload)
        # rans_decode_fea(decode_rounded_min=False), model.py:377-393
        clen = np.empty(1, dtype=np.uint16)
        dec.decode(self.cdf2, clen)
        cdf = np.empty(int(clen[0]) + 1, dtype=np.uint16)
        dec.decode(self.cdf1, cdf)
        cdf = np.pad(cdf + np.uint16(1), (0, 1))
        cdf[-1] = 65535
        bottom = np.empty(nb * 3, dtype=np.uint16)
        dec.decode(cdf[None], bottom)
        L = self.n_ds - self.skip
        c0 = np.concatenate([np.zeros((nb, 1), np.int32), bottom.astype(np.int32).reshape(-1, 3)], 1)
        cur = SparseTensor(np.ones((nb, 1), dtype=np.int8), c0, 2 ** L)

        def decode_oct(logits):
            q = K.batch_quantize_pmf(logits)
            out = np.empty(q.shape[0], dtype=np.uint16)
            dec.decode(q, out)
            return out

        bins, top_rec, top_stride, cur_bin = [], None, None, None
        for idx, blk in self._levels():
            if isinstance(blk, OneScalePredictor):
                cur, cur_bin = blk.decompress(cur, decode_oct, idx != 1 and blk.if_upsample)
            else:
                bins.append(cur_bin)
                cur, cur_bin, top_rec, top_stride = blk.decompress(cur, bins, top_rec, top_stride, decode_oct)
        base = cur.C if top_rec is None else top_rec
        assert (cur.stride[0] if top_rec is None else top_stride) == 2
        rec = base[:, None, 1:].copy()
        rec <<= 1
        rec = (rec + UNFOLD[:, :, 1:])[cur_bin]
        return rec + off[None]
