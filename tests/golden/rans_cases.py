"""Known-answer cases for the range coders (SURVEY.md 8c KAT-A..I + the reference's import-time
self tests).  `run_all` replays every case against ANY implementation with the reference's class
surface and returns {case: hex | sha256 | values}; the same function is used to mint the golden
file from the compiled reference and to check the oracle and the CUDA coders against it."""
import hashlib

import numpy as np


def _h(b: bytes):
    return {'len': len(b), 'sha256': hashlib.sha256(b).hexdigest()} if len(b) > 64 else b.hex()


QUAN_CDF = np.array([[1, 2, 3, 4, 65535], [1, 2, 3, 5, 65535], [2, 3, 4, 6, 65535],
                     [2, 3, 4, 7, 65535], [1, 2, 3, 8, 65535], [1, 2, 3, 9, 65535]], dtype=np.uint16)
QUAN_CDF2 = np.array([[1, 2, 4000, 5000, 65535], [2, 3, 3000, 6000, 65535], [3, 4, 3000, 7000, 65535],
                      [4, 5, 1000, 8000, 65535], [5, 6, 5000, 9000, 65535], [6, 7, 6000, 10000, 65535]],
                     dtype=np.uint16)
ORG = np.array([2, 4, 1, 1, 2, 3, 0, 2, 4, 2, 1, 1], dtype=np.uint16)


def kat_e_inputs(n=100000):
    rng = np.random.default_rng(1234)
    pm = rng.integers(1, 200, (n, 255))
    pm = pm * (65536 - 255) // pm.sum(1, keepdims=True) + 1
    cdf = np.cumsum(pm, 1)
    cdf[:, -1] = 65535
    sym = rng.integers(0, 255, n)
    return cdf.astype(np.uint16), sym.astype(np.uint16)


def kat_gh_inputs():
    rng = np.random.default_rng(4321)
    n = 200000
    p = np.clip(np.round(rng.random((1, n)) ** 3 * 65536), 1, 65535).astype(np.uint32)
    s = (rng.random((1, n)) * 65536 < p)
    t = rng.integers(-20, 21, (50000, 1)).astype(np.int32)
    t[rng.random(50000) < 0.7] = 0
    return p, s, t


def run_all(RansEncoder, RansDecoder, IndexedRansCoder, BinaryRansCoder, batched_pmf_to_quantized_cdf):
    out = {}
    # ---- simple coder: lossy_coord_v3/rans_coder/__init__.py:28-63 ----
    enc = RansEncoder(8 * 1024 * 1024)
    out['A_n1'] = int(enc.encode(QUAN_CDF2, ORG[6:12]))
    out['A_n2'] = int(enc.encode(QUAN_CDF, ORG[:6]))
    s = enc.flush()
    out['A'] = _h(s)
    dec = RansDecoder()
    dec.flush(s)
    d = np.zeros(12, dtype=np.uint16)
    dec.decode(QUAN_CDF, d[:6]); dec.decode(QUAN_CDF2, d[6:12])
    out['A_roundtrip'] = bool((d == ORG).all())
    enc.encode(QUAN_CDF, ORG[:6]); enc.encode(QUAN_CDF2, ORG[6:12])
    s = enc.flush()
    out['B'] = _h(s)
    dec.flush(s)
    d = np.zeros(12, dtype=np.uint16)
    dec.decode(QUAN_CDF2, d[6:12]); dec.decode(QUAN_CDF, d[:6])
    out['B_roundtrip'] = bool((d == ORG).all())
    enc.encode(QUAN_CDF, ORG[:6])
    out['A_single'] = _h(enc.flush())
    # shared side-info CDF (lossl_coord_int/model.py:255-257)
    cdf2 = (np.arange(1, 129, dtype=np.uint16)[None] * 512)
    cdf2[:, -1] = 65535
    enc.encode(cdf2, np.array([0, 5, 127, 64], dtype=np.uint16))
    s = enc.flush()
    out['C'] = _h(s)
    dec.flush(s)
    d = np.zeros(4, dtype=np.uint16); dec.decode(cdf2, d)
    out['C_dec'] = d.tolist()
    out['D_empty'] = _h(enc.flush())
    cdf, sym = kat_e_inputs()
    enc32 = RansEncoder(32 * 1024 * 1024)
    enc32.encode(cdf, sym)
    s = enc32.flush()
    out['E'] = _h(s)
    dec.flush(s)
    d = np.zeros(sym.shape[0], dtype=np.uint16); dec.decode(cdf, d)
    out['E_roundtrip'] = bool((d == sym).all())
    # binary mode of the simple coder
    rng = np.random.default_rng(77)
    cb = rng.integers(1, 65535, 5000).astype(np.uint16)
    sb = rng.random(5000) < cb / 65536.0
    enc.encode_bin(cb, sb)
    s = enc.flush()
    out['simple_bin'] = _h(s)
    dec.flush(s)
    d = np.zeros(5000, dtype=np.bool_); dec.decode_bin(cb, d)
    out['simple_bin_roundtrip'] = bool((d == sb).all())

    # ---- cdf_ops KAT: lib/entropy_models/rans_coder/__init__.py:72-77 ----
    pm = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1], [2 ** -17, 1, 0, 0]], dtype=np.float64)
    off = np.zeros(4, dtype=np.int32)
    coder = IndexedRansCoder(True, 8, 100)
    coder.init_with_pmfs(pm.copy(), off)
    out['cdf_kat'] = [list(map(int, c)) for c in coder.get_cdfs()]
    out['cdf_kat_offsets'] = off.tolist()
    sym = np.array([[0], [1], [0], [1], [0], [1], [3], [3]], dtype=np.int32)
    idx = np.array([[0], [0], [1], [1], [2], [2], [3], [3]], dtype=np.int32)
    enc_list = coder.encode_with_indexes(sym, idx)
    out['indexed8'] = [bytes(b).hex() for b in enc_list]
    d = np.empty_like(sym); coder.decode_with_indexes(enc_list, idx, d)
    out['indexed8_roundtrip'] = bool((d == sym).all())
    # random pmfs through the free function, both modes
    rng = np.random.default_rng(5)
    pm = rng.random((6, 9)) ** 4
    pm[2, 3:] = 0; pm[4, :4] = 0; pm[5] = pm[5] / pm[5].sum() * 0.6
    for mode in (True, False):
        off = np.arange(-3, 3, dtype=np.int32)
        cd = batched_pmf_to_quantized_cdf(pm.copy(), off, mode)
        out[f'cdfs_random_overflow{int(mode)}'] = [list(map(int, c)) for c in cd]
        out[f'cdfs_random_overflow{int(mode)}_offsets'] = off.tolist()

    # ---- binary coder (KAT-F, KAT-G) ----
    bc = BinaryRansCoder(1, 100)
    sF = np.array([[1, 0, 0, 1, 1, 1, 0, 1]], dtype=np.bool_)
    pF = np.array([[40000, 100, 65535, 1, 32768, 60000, 5, 12345]], dtype=np.uint32)
    out['F'] = bytes(bc.encode(sF, pF)[0]).hex()
    p, s, t = kat_gh_inputs()
    eb = bc.encode(s, p)
    out['G'] = _h(bytes(eb[0]))
    d = np.empty_like(s); bc.decode(eb, p, d)
    out['G_roundtrip'] = bool((d == s).all())
    # ---- KAT-H: rans_encode_with_cdf path (geo_lossl_em.py:59-74) ----
    tmin = int(t.min())
    pmf = np.bincount((t - tmin).reshape(-1)).astype(np.float64)
    ic = IndexedRansCoder(False, 1)
    off = np.array([tmin], dtype=np.int32)
    ic.init_with_pmfs(pmf[None].copy(), off)
    out['H_cdf'] = list(map(int, ic.get_cdfs()[0]))
    eh = ic.encode(t.reshape(1, -1))
    out['H'] = _h(bytes(eh[0]))
    d = np.empty((1, t.shape[0]), dtype=np.int32); ic.decode(eh, d)
    out['H_roundtrip'] = bool((d.reshape(-1) == t.reshape(-1)).all())
    # ---- KAT-I: overflow coding ----
    oc = IndexedRansCoder(True, 1)
    off = np.array([-2], dtype=np.int32)
    oc.init_with_pmfs(np.array([[.1, .2, .4, .2, .1]], dtype=np.float64), off)
    out['I_cdf'] = list(map(int, oc.get_cdfs()[0]))
    out['I_offset'] = off.tolist()
    sI = np.array([[0, -2, 2, -3, 3, 100, -100, 1, -1, 0]], dtype=np.int32)
    eI = oc.encode(sI)
    out['I'] = bytes(eI[0]).hex()
    d = np.empty_like(sI); oc.decode(eI, d)
    out['I_roundtrip'] = bool((d == sI).all())
    # overflow + indexes + batch, larger
    rng = np.random.default_rng(99)
    pm = rng.random((5, 12)) + 0.01
    pm /= pm.sum(1, keepdims=True) * 1.02
    off = np.full(5, -6, dtype=np.int32)
    mc = IndexedRansCoder(True, 3)
    mc.init_with_pmfs(pm.copy(), off)
    symM = np.clip(np.round(rng.normal(0, 6, (3, 4000))), -3000, 3000).astype(np.int32)
    symM[0, :5] = [-3000, 3000, 2049, -2049, 0]
    idxM = rng.integers(0, 5, (3, 4000)).astype(np.int32)
    eM = mc.encode_with_indexes(symM, idxM)
    out['M'] = [_h(bytes(b)) for b in eM]
    d = np.empty_like(symM); mc.decode_with_indexes(eM, idxM, d)
    out['M_roundtrip'] = bool((d == symM).all())
    eM2 = mc.encode(symM)
    out['M_noidx'] = [_h(bytes(b)) for b in eM2]
    return out
