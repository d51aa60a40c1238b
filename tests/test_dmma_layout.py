"""Host-side model of the shared-memory layouts of the producer-warp FP64 tensor-core kernels
(c-attl3_b200/csrc/conv_dmma.cu, dmma2_*): the index arithmetic restated in numpy and checked for the properties the
kernels rely on -- every element of a stage written exactly once by the producers, fragment reads free of bank
conflicts, the packed-weight order equal to the order the gather GEMM consumes.  No GPU needed; the kernels themselves
are checked against the oracle by tests/test_gpu_parity.py::test_conv_double_producer_warp_kernels_vs_oracle."""
import itertools

import numpy as np
import pytest

BK, PAD, PRODUCERS = 8, 4, 128          # D2_BK, DM_PAD, D2_PRODUCERS
TILES = [(2, 4, 4), (4, 2, 4), (1, 8, 4)]   # (WR, WC, NI): 128 x 128, 256 x 64, 64 x 256


def sw_offset(row, kk):
    """Element (row, kk) of a weight-gradient tile [row][8 m]: the 16-byte chunk index XOR bit 1 of the row."""
    return row * BK + 2 * ((kk >> 1) ^ (row & 2)) + (kk & 1)


def test_swizzle_is_a_permutation_of_each_row():
    for row in range(16):
        offs = sorted(sw_offset(row, kk) - row * BK for kk in range(BK))
        assert offs == list(range(BK))


@pytest.mark.parametrize("k4", [0, 4])
def test_swizzled_fragment_reads_hit_sixteen_banks_per_half_warp(k4):
    # lane -> (g = lane / 4, kq = lane % 4) reads element (row0 + 8 mi + g, k4 + kq); an 8-byte access of a half warp is
    # conflict free when its 16 addresses fall into 16 different 8-byte banks (16 x 8 B = the 128-byte bank line)
    for base in (0, 8, 64):
        for half in (0, 1):
            banks = set()
            for lane in range(16 * half, 16 * half + 16):
                g, kq = lane >> 2, lane & 3
                banks.add(sw_offset(base + g, k4 + kq) % 16)
            assert len(banks) == 16


@pytest.mark.parametrize("tile", [128, 256, 64])
@pytest.mark.parametrize("k4", [0, 4])
def test_padded_fragment_reads_hit_sixteen_banks_per_half_warp(tile, k4):
    # gather GEMM tiles are [k][row] with a pitch of tile + 4 doubles
    pitch = tile + PAD
    for half in (0, 1):
        banks = set()
        for lane in range(16 * half, 16 * half + 16):
            g, kq = lane >> 2, lane & 3
            banks.add(((k4 + kq) * pitch + g) % 16)
        assert len(banks) == 16


@pytest.mark.parametrize("wr,wc,ni", TILES)
def test_weight_gradient_producers_fill_a_stage_exactly_once(wr, wc, ni):
    rows_a, rows_b = 64 * wr, 8 * ni * wc
    for rows in (rows_a, rows_b):
        hits = np.zeros(rows * BK, dtype=int)
        for pt in range(PRODUCERS):
            ch, rg = pt & 3, pt >> 2
            dst_off = 2 * (ch ^ (rg & 2))
            for i in range(rows // 32):
                row = rg + 32 * i
                assert row & 2 == rg & 2
                base = row * BK + dst_off
                assert base % 2 == 0                     # 16-byte aligned destination
                hits[base:base + 2] += 1
                # the pair holds m offsets 2 ch, 2 ch + 1 of this row, where the consumers look for them
                assert base == sw_offset(row, 2 * ch) and base + 1 == sw_offset(row, 2 * ch + 1)
        assert (hits == 1).all()


@pytest.mark.parametrize("wr,wc,ni", TILES[:2])
@pytest.mark.parametrize("vec", [True, False])
def test_gather_producers_fill_a_stage_exactly_once(wr, wc, ni, vec):
    bm, bn = 64 * wr, 8 * ni * wc
    pa, pb = bm + PAD, bn + PAD
    a = np.zeros(BK * pa, dtype=int)
    pairs = bm // 2
    ksplit = PRODUCERS // pairs if vec else 1
    kper = BK // ksplit
    for pt in range(PRODUCERS):
        if vec:
            rows, k_first, width = [2 * (pt % pairs)], (pt // pairs) * kper, 2
        else:
            rows, k_first, width = [pt + PRODUCERS * u for u in range(bm // PRODUCERS)], 0, 1
        for row, kk in itertools.product(rows, range(kper)):
            off = (k_first + kk) * pa + row
            if vec:
                assert off % 2 == 0
            a[off:off + width] += 1
    for k in range(BK):
        assert (a[k * pa:k * pa + bm] == 1).all() and (a[k * pa + bm:(k + 1) * pa] == 0).all()
    b = np.zeros(BK * pb, dtype=int)
    for pt in range(PRODUCERS):
        for i in range(BK * bn // 2 // PRODUCERS):
            idx = pt + PRODUCERS * i
            k, cp = divmod(idx, bn // 2)
            b[k * pb + 2 * cp:k * pb + 2 * cp + 2] += 1
    for k in range(BK):
        assert (b[k * pb:k * pb + bn] == 1).all() and (b[k * pb + bn:(k + 1) * pb] == 0).all()


@pytest.mark.parametrize("rh,rw,sc,j,bn", [(3, 3, 20, 40, 64), (1, 1, 64, 256, 128), (5, 5, 9, 48, 64), (3, 2, 7, 130, 64)])
def test_packed_weights_follow_the_consumption_order(rh, rw, sc, j, bn):
    # dmma2_pack_weights_kernel: [column tile][k-step][8 reduce rows][BN columns], k-step = channel block * taps + tap, taps
    # numbered rh + RH * rw; the producers walk rh fastest, then rw, then the channel block
    taps = rh * rw
    rblocks = -(-sc // BK)
    ksteps = taps * rblocks
    j_tiles = -(-j // bn)
    w = np.arange(taps * sc * j, dtype=np.int64).reshape(taps, sc, j) + 1     # w[tap][r][j], all non-zero
    total = j_tiles * ksteps * BK * bn
    wp = np.zeros(total, dtype=np.int64)
    for i in range(total):
        col = i % bn
        t = i // bn
        k = t % BK
        u = t // BK
        ks, jt = u % ksteps, u // ksteps
        rb, tap = divmod(ks, taps)
        r, jj = rb * BK + k, jt * bn + col
        wp[i] = w[tap, r, jj] if r < sc and jj < j else 0
    ks = 0
    for r0 in range(0, rblocks * BK, BK):
        for rw_i in range(rw):
            for rh_i in range(rh):
                tap = rh_i + rh * rw_i
                for jt in range(j_tiles):
                    tile = wp[(jt * ksteps + ks) * BK * bn:(jt * ksteps + ks + 1) * BK * bn].reshape(BK, bn)
                    for k in range(BK):
                        for col in (0, 1, bn - 1):
                            r, jj = r0 + k, jt * bn + col
                            want = w[tap, r, jj] if r < sc and jj < j else 0
                            assert tile[k, col] == want
                ks += 1
    assert ks == ksteps
    assert np.count_nonzero(wp) == w.size      # every weight appears exactly once
