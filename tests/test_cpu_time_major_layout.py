"""CPU tier: host-side decisions and shared-memory layouts of the time-major conv kernel (csrc/conv_tc.cu conv1d_tct_kernel),
restated in Python: which layers take it, the XOR-swizzled residual / value slot (a bijection, lane-private words, bank-
conflict-free in both views), and the two statistics formulas of its epilogue against a float64 variance."""
import math

import numpy as np
import pytest


def test_time_major_mode_selection(monkeypatch):
    """FAST-recipe convs with Cout <= min(ST2_TC_TMAJOR_MAX, 128) carry the ST2_TC_TMAJOR flag; ACCURATE / F16X3 never do."""
    from styletts2_b200 import ops
    from styletts2_b200.lib import TC_ACCURATE, TC_F16X3, TC_FAST, TC_TMAJOR
    monkeypatch.setattr(ops, "TC_MODE_OVERRIDE", None)
    monkeypatch.setattr(ops, "TC_TMAJOR_MAX_COUT", 128)
    for co in (1, 16, 22, 32, 64, 128):
        assert ops._tc_mode(TC_FAST, co) == TC_FAST | TC_TMAJOR
    for co in (129, 256, 1024):
        assert ops._tc_mode(TC_FAST, co) == TC_FAST
    assert ops._tc_mode(TC_ACCURATE, 64) == TC_ACCURATE and ops._tc_mode(TC_F16X3, 64) == TC_F16X3
    assert ops._tc_mode(TC_FAST) == TC_FAST                      # no channel count given: plain recipe
    monkeypatch.setattr(ops, "TC_TMAJOR_MAX_COUT", 64)
    assert ops._tc_mode(TC_FAST, 64) & TC_TMAJOR and not ops._tc_mode(TC_FAST, 128) & TC_TMAJOR
    monkeypatch.setattr(ops, "TC_TMAJOR_MAX_COUT", 0)
    assert ops._tc_mode(TC_FAST, 32) == TC_FAST
    monkeypatch.setattr(ops, "TC_TMAJOR_MAX_COUT", 4096)
    assert not ops._tc_mode(TC_FAST, 256) & TC_TMAJOR            # 128 is the kernel's limit whatever the variable says


def _word_addr(j, lane):
    """byte offset of (channel j, frame `lane`) in a 2 KB slot: ((slot + lane_off) ^ ((j & 7) << 4)) + 128 j with slot = 0"""
    lane_off = ((lane >> 2) << 4) | ((lane & 3) << 2)
    return (lane_off ^ ((j & 7) << 4)) + 128 * j


def test_swizzled_slot_is_a_bijection_and_conflict_free():
    words = {}
    for j in range(16):
        row = [_word_addr(j, lane) for lane in range(32)]
        # one warp instruction = the 32 frames of one channel: stays inside the channel's 128-byte row, all 32 banks distinct
        assert all(128 * j <= a < 128 * (j + 1) for a in row)
        assert len({(a >> 2) & 31 for a in row}) == 32
        for lane, a in enumerate(row):
            words[a] = (j, lane)
    assert len(words) == 16 * 32 and min(words) == 0 and max(words) == 2048 - 4
    # transposed view of the statistics: lane -> (channel lane / 2, frames 16 * (lane & 1) .. + 15) reads four 16-byte chunks
    # at (t_off ^ (k << 4)); every chunk must hold exactly the four consecutive frames it stands for, and the eight lanes of a
    # 128-bit shared-memory phase must hit eight different 16-byte bank groups
    for k in range(4):
        chunks = []
        for lane in range(32):
            sch, shalf = lane >> 1, lane & 1
            t_off = sch * 128 + (((shalf << 2) ^ (sch & 7)) << 4)
            a = t_off ^ (k << 4)
            frames = [words[a + 4 * w] for w in range(4)]
            assert frames == [(sch, 16 * shalf + 4 * k + w) for w in range(4)]
            chunks.append(a)
        for phase in range(4):
            assert len({(a >> 4) & 7 for a in chunks[8 * phase:8 * phase + 8]}) == 8


def test_fp8_correction_k_order_pairs_the_right_planes():
    """Correction MMA (K = 32 e4m3): activation chunk c = [h(z) of channels 8c..8c+7 | l(z) of the same], weight chunk c =
    [l(w) | h(w)] of the same channels: the dot product over the 32 K elements is sum_c h(z_c) l(w_c) + l(z_c) h(w_c)."""
    rng = np.random.default_rng(0)
    hz, lz, hw, lw = (rng.standard_normal(16) for _ in range(4))
    a_row = np.concatenate([np.concatenate([hz[8 * c:8 * c + 8], lz[8 * c:8 * c + 8]]) for c in range(2)])
    w_row = np.concatenate([np.concatenate([lw[8 * c:8 * c + 8], hw[8 * c:8 * c + 8]]) for c in range(2)])
    assert math.isclose(float(a_row @ w_row), float(hz @ lw + lz @ hw), rel_tol=1e-12)


@pytest.mark.parametrize("mean,spread", [(0.0, 1.0), (10.0, 0.01), (-300.0, 2.0)])
def test_epilogue_statistics_formulas(mean, spread):
    """Full steps: two 16-frame halves (two-pass mean / M2 in fp32) merged with Chan's formula; partial steps: one pass over the
    deviations from a pilot sample.  Both must reproduce the float64 (count, mean, M2) of the 32 (or fewer) values."""
    rng = np.random.default_rng(1)
    x = (mean + spread * rng.standard_normal(32)).astype(np.float32)
    ref_mean, ref_m2 = float(x.astype(np.float64).mean()), float(((x.astype(np.float64) - x.astype(np.float64).mean()) ** 2).sum())
    # full-step path
    halves = []
    for h in range(2):
        xs = x[16 * h:16 * h + 16]
        mh = np.float32(xs.sum(dtype=np.float32) * np.float32(1 / 16))
        qh = np.float32(((xs - mh) ** 2).sum(dtype=np.float32))
        halves.append((mh, qh))
    dl = np.float32(halves[1][0] - halves[0][0])
    m_full = np.float32(halves[0][0] + np.float32(0.5) * dl)
    q_full = np.float32(halves[0][1] + halves[1][1] + dl * dl * np.float32(8))
    assert abs(float(m_full) - ref_mean) <= 2e-6 * max(1.0, abs(ref_mean))
    assert abs(float(q_full) - ref_m2) <= 2e-5 * ref_m2 + 1e-10
    # partial-step path (nvalid = 21 frames)
    nv = 21
    xv = x[:nv]
    d = (xv - xv[0]).astype(np.float32)
    s1, s2 = np.float32(d.sum(dtype=np.float32)), np.float32((d * d).sum(dtype=np.float32))
    m_part = np.float32(xv[0] + s1 / np.float32(nv))
    q_part = np.float32(max(0.0, float(s2 - s1 * s1 / np.float32(nv))))
    rm = float(xv.astype(np.float64).mean())
    rq = float(((xv.astype(np.float64) - rm) ** 2).sum())
    assert abs(float(m_part) - rm) <= 2e-6 * max(1.0, abs(rm))
    assert abs(float(q_part) - rq) <= 1e-4 * rq + 1e-10
