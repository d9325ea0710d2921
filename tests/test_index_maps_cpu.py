"""Integer index maps of the CUDA kernels, restated in Python and checked BIT-EXACTLY on the CPU against the tensor
operations of the reference path (torch.roll + HF window_partition, the shift-mask construction of ScOTLayer.get_attn_mask,
HF's relative_position_index, the patch-merging slice order, the patch-unmerging permute).

Each `*_formula` function below is a literal transcription of the device code it names (same variable names, same
integer arithmetic); the GPU tests check the kernels end to end, this file pins the index arithmetic itself — the part
of the path that must be exact (BASELINE.json north_star: "bit-exact for window-partition/index ops").
"""
import itertools

import pytest
import torch

from oracle import scot_oracle as O


# ---- attention.cu: token_row<WS>() -------------------------------------------------------------------------------
def token_row_formula(res, ws, shift, bw, n):
    nws = res // ws
    nw = nws * nws
    b, w = divmod(bw, nw)
    wi, wj = divmod(w, nws)
    i = wi * ws + n // ws + shift
    j = wj * ws + n % ws + shift
    if i >= res:
        i -= res
    if j >= res:
        j -= res
    return (b * res + i) * res + j


@pytest.mark.parametrize("res,ws,shift", [(32, 16, 0), (32, 16, 8), (16, 8, 4), (64, 16, 8), (16, 16, 0), (8, 8, 0), (4, 4, 0)])
def test_token_row_equals_roll_plus_window_partition(res, ws, shift):
    """reference: shifted = torch.roll(x, (-s, -s), (1, 2)); windows = window_partition(shifted, ws)  (model.py:522-531)"""
    B = 2
    ids = torch.arange(B * res * res).view(B, res, res, 1)
    x = torch.roll(ids, (-shift, -shift), (1, 2)) if shift else ids
    win = O.window_partition(x, ws).view(-1, ws * ws)  # [B * nW, N] -> original flat token row of every window slot
    for bw in range(win.shape[0]):
        got = [token_row_formula(res, ws, shift, bw, n) for n in range(ws * ws)]
        assert got == win[bw].tolist()


# ---- attention.cu: win_flags() / mask_code<WS>() -------------------------------------------------------------------
def mask_code_formula(res, ws, shift, bw, n):
    nws = res // ws
    w = bw % (nws * nws)
    wi, wj = divmod(w, nws)
    flags = 0 if shift == 0 else ((1 if wi == nws - 1 else 0) | (2 if wj == nws - 1 else 0))
    hm = 1 if (flags & 1) and (n // ws >= ws - shift) else 0
    wm = 1 if (flags & 2) and (n % ws >= ws - shift) else 0
    return hm | (wm << 1)


@pytest.mark.parametrize("res,ws,shift", [(32, 16, 8), (16, 8, 4), (64, 16, 8)])
def test_mask_code_equals_reference_shift_mask(res, ws, shift):
    """reference: img_mask regions 0..8 over the slices (0,-ws), (-ws,-shift), (-shift,None); mask[i,j] = -100 where the
    region ids of the two tokens of a window differ (model.py:442-478). The kernels compare 2-bit region codes."""
    img = torch.zeros(1, res, res, 1)
    cnt = 0
    for hs, wsl in itertools.product((slice(0, -ws), slice(-ws, -shift), slice(-shift, None)), repeat=2):
        img[:, hs, wsl, :] = cnt
        cnt += 1
    mw = O.window_partition(img, ws).view(-1, ws * ws)
    ref = (mw.unsqueeze(1) - mw.unsqueeze(2)) != 0  # [nW, N, N]
    N = ws * ws
    for bw in range(ref.shape[0]):
        code = torch.tensor([mask_code_formula(res, ws, shift, bw, n) for n in range(N)])
        got = code.unsqueeze(0) != code.unsqueeze(1)
        assert torch.equal(got, ref[bw])


# ---- attention.cu: bias_rowbase<WS>() / bias_coloff<WS>() ------------------------------------------------------------
def rel_index_formula(ws, m, n):
    rowbase = (m // ws) * (2 * ws - 1) + (m % ws) + (ws - 1) * (2 * ws - 1) + (ws - 1)
    coloff = (n // ws) * (2 * ws - 1) + (n % ws)
    return rowbase - coloff


@pytest.mark.parametrize("ws", [4, 8, 16])
def test_bias_lookup_equals_hf_relative_position_index(ws):
    """HF Swinv2SelfAttention.create_coords_table_and_index (modeling_swinv2.py:512-523)"""
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    ref = rel.sum(-1)
    N = ws * ws
    got = torch.tensor([[rel_index_formula(ws, m, n) for n in range(N)] for m in range(N)])
    assert torch.equal(got, ref)


# ---- attention.cu: fold_bias_to_table<WS, NWARP>() vs fold_bias_to_table16<NWARP>() ----------------------------------
def fold_slots(ws, nwarp, rg, r):
    N, MT, SIDE = ws * ws, ws * ws // 16, 2 * ws - 1
    WPI, ACC = max(nwarp // MT, 1), 16 * N
    mt_lo = rg * nwarp if MT > nwarp else 0
    mt_hi = mt_lo + nwarp if MT > nwarp else MT
    dp, dq = r // SIDE - (ws - 1), r % SIDE - (ws - 1)
    pm0, pm1 = max(dp, 0), (ws + dp if dp < 0 else ws)
    qm0, qm1 = max(dq, 0), (ws + dq if dq < 0 else ws)
    if MT > nwarp:
        pm0, pm1 = max(pm0, mt_lo), min(pm1, mt_hi)
    out = []
    for pm in range(pm0, pm1):
        for qm in range(qm0, qm1):
            m, n = pm * ws + qm, (pm - dp) * ws + (qm - dq)
            lane = ((m & 7) << 2) | ((n & 7) >> 1)
            reg = (((m >> 3) & 1) << 1) | (n & 1)
            slot = (((n >> 3) << 2) + reg) * 32 + lane
            out += [((c * MT + (m >> 4) - mt_lo) * ACC + slot, m, n) for c in range(WPI)]
    return out


def fold16_slots(nwarp, rg, r):
    ws, N, MT, SIDE = 16, 256, 16, 31
    ACC = 16 * N
    mt_lo, mt_hi = rg * nwarp, rg * nwarp + nwarp
    dp, dq = r // SIDE - (ws - 1), r % SIDE - (ws - 1)
    pm0, pm1 = max(dp, 0, mt_lo), min((ws + dp if dp < 0 else ws), mt_hi)
    qm0, qm1 = max(dq, 0), (ws + dq if dq < 0 else ws)
    d = dp * ws + dq
    out = []
    for qm in range(qm0, qm1):
        m = pm0 * ws + qm
        n = m - d
        p = ((m >> 4) - mt_lo) * ACC + ((m >> 3) & 1) * 64 + (m & 7) * 4 + (n >> 3) * 128 + (n & 1) * 32 + ((n & 7) >> 1)
        for _pm in range(pm0, pm1):
            out.append(p)
            p += ACC + 256
    return out


@pytest.mark.parametrize("ws,nwarp", [(16, 8), (16, 4), (8, 8), (4, 8)])
def test_bias_gradient_fold_covers_every_pair_once(ws, nwarp):
    """Each (query m, key n) pair of a window is read exactly once, by the table entry of its displacement, from the slot
    the mma fragment layout puts it in (m16n8 accumulator: lane = (m%8)*4 + (n%8)/2, reg = ((m%16)/8)*2 + n%2)."""
    N, MT, SIDE = ws * ws, ws * ws // 16, 2 * ws - 1
    seen = set()
    for rg in range(max(MT // nwarp, 1)):
        for r in range(SIDE * SIDE):
            for _addr, m, n in fold_slots(ws, nwarp, rg, r):
                assert (m // ws - n // ws + ws - 1) * SIDE + (m % ws - n % ws + ws - 1) == r  # its own table entry
                seen.add((m, n, _addr // (16 * N) if MT <= nwarp else 0))
    pairs = {(m, n) for m, n, _ in seen}
    assert len(pairs) == N * N
    # the table entry r = (dp + ws-1) * SIDE + (dq + ws-1) with dp = p_m - p_n, dq = q_m - q_n is HF's relative index
    for m, n in [(0, 0), (N - 1, 0), (0, N - 1), (17 % N, 5 % N)]:
        dp, dq = m // ws - n // ws, m % ws - n % ws
        assert (dp + ws - 1) * SIDE + (dq + ws - 1) == rel_index_formula(ws, m, n)


@pytest.mark.parametrize("nwarp", [8, 4])
def test_fold16_reads_the_same_slots(nwarp):
    for rg in range(16 // nwarp):
        for r in range(31 * 31):
            assert sorted(a for a, _, _ in fold_slots(16, nwarp, rg, r)) == sorted(fold16_slots(nwarp, rg, r))


# ---- misc.cu: merge_gather (patch merging) and norm.cu: unmerge_row() ----------------------------------------------
def test_merge_order_equals_reference_slices():
    """ScOTPatchMerging: cat([x[0::2,0::2], x[1::2,0::2], x[0::2,1::2], x[1::2,1::2]], -1)  (model.py:694-704);
    gathered row (b, i, j) holds the channels of tokens (2i+di, 2j+dj) in the order (0,0), (1,0), (0,1), (1,1)."""
    B, res, C = 2, 8, 3
    x = torch.arange(B * res * res * C, dtype=torch.float32).view(B, res, res, C)
    ref = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1).reshape(-1, 4 * C)
    got = torch.empty_like(ref)
    h = res // 2
    for b in range(B):
        for i in range(h):
            for j in range(h):
                for q, (di, dj) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
                    got[(b * h + i) * h + j, q * C:(q + 1) * C] = x[b, 2 * i + di, 2 * j + dj]
    assert torch.equal(got, ref)


def unmerge_row_formula(r_in, res):
    ac, m = r_in & 3, r_in >> 2
    j, t = m % res, m // res
    i, b = t % res, t // res
    a, c = ac >> 1, ac & 1
    return (b * (2 * res) + 2 * i + a) * (2 * res) + 2 * j + c


@pytest.mark.parametrize("res", [2, 4, 8])
def test_unmerge_row_equals_reference_permute(res):
    """ScOTPatchUnmerging: reshape [B,h,w,2,2,C/2] -> permute(0,1,3,2,4,5) -> [B, 4hw, C/2]  (model.py:748-754)"""
    B, Ch = 2, 1
    z = torch.arange(B * res * res * 4 * Ch).view(B * res * res, 4 * Ch)
    ref = z.reshape(B, res, res, 2, 2, Ch).permute(0, 1, 3, 2, 4, 5).reshape(-1)
    rows = B * res * res * 4
    got = torch.empty(rows, dtype=ref.dtype)
    for r_in in range(rows):
        got[unmerge_row_formula(r_in, res)] = z.view(-1)[r_in]
    assert torch.equal(got, ref)
