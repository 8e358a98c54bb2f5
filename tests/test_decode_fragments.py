"""CPU: the operand maps of the decode kernels' tensor-core products (tests/_decode_fragments.py, the lane-level model
gsr_decode.cu follows expression by expression) reproduce the plain matrix products of the four MLPs
(scene/gaussian_model.py:118-144), forward and backward, for every supported n_offsets class."""
import numpy as np
import pytest

import _decode_fragments as fr


def _weights(k, rng):
    w1 = [rng.standard_normal((32, 36)) for _ in range(4)]
    b1 = [rng.standard_normal(32) for _ in range(4)]
    w2 = [rng.standard_normal((fr.out_count(m, k), 32)) for m in range(4)]
    b2 = [rng.standard_normal(fr.out_count(m, k)) for m in range(4)]
    return w1, b1, w2, b2


def _b2pad(b2, k):
    cb, cols, S = fr.col_plan(k)
    out = np.zeros(cols + 8)
    for m in range(4):
        out[cb[m]:cb[m] + fr.out_count(m, k)] = b2[m]
    return out


def _c_tile(frag):
    C = np.zeros((16, 8))
    C[fr.G, 2 * fr.T] = frag[0]; C[fr.G, 2 * fr.T + 1] = frag[1]; C[fr.G + 8, 2 * fr.T] = frag[2]; C[fr.G + 8, 2 * fr.T + 1] = frag[3]
    return C


@pytest.mark.parametrize("k", [1, 5, 10, 11, 16])
def test_part_a_products(k):
    rng = np.random.default_rng(k)
    w1, b1, w2, b2 = _weights(k, rng)
    cb, cols, S = fr.col_plan(k)
    assert S % 8 == 4 and (S // 4) % 2 == 1 and cols % 16 == 0
    x = rng.standard_normal((16, 36))
    W1F, W1B, W2F, W2B = fr.build_W1F(w1), fr.build_W1B(w1), fr.build_W2F(w2, k), fr.build_W2B(w2, k)
    b2p = _b2pad(b2, k)
    xa = fr.x_fragments(x)
    OUT = np.zeros(16 * S)
    hs = []
    for m in range(4):
        h = fr.layer1_forward(xa, W1F, b1, m)
        H = np.concatenate([_c_tile(t) for t in h], axis=1)
        ref = np.maximum(x @ w1[m].T + b1[m], 0.0)
        np.testing.assert_allclose(H, ref, atol=1e-12)
        hs.append(ref)
        fr.layer2_forward(h, W2F, b2p, k, m, OUT, S)
    O = OUT.reshape(16, S)
    for m in range(4):
        n = fr.out_count(m, k)
        np.testing.assert_allclose(O[:, cb[m]:cb[m] + n], hs[m] @ w2[m].T + b2[m], atol=1e-12)
        pad_end = (cb[m + 1] if m < 3 else cols)
        assert np.all(O[:, cb[m] + n:pad_end] == 0.0)     # pad columns stay finite (zero)

    # backward: random gradient of the pre-activations in the valid columns
    D = np.zeros((16, S))
    for m in range(4):
        D[:, cb[m]:cb[m] + fr.out_count(m, k)] = rng.standard_normal((16, fr.out_count(m, k)))
    Dflat = D.reshape(-1)
    dx = [[np.zeros(32) for _ in range(4)] for _ in range(5)]
    dx_ref = np.zeros((16, 36))
    for m in range(4):
        dh = fr.layer2_backward(Dflat, S, W2B, k, m)
        dH = np.concatenate([_c_tile(t) for t in dh], axis=1)
        ref = D[:, cb[m]:cb[m] + fr.out_count(m, k)] @ w2[m]
        np.testing.assert_allclose(dH, ref, atol=1e-12)
        gate = hs[m] > 0
        dh = [[np.where(_gate_frag(gate, nt, e), dh[nt][e], 0.0) for e in range(4)] for nt in range(4)]
        dx = fr.layer1_backward(dx, dh, W1B, m)
        dx_ref += (ref * gate) @ w1[m]
    # lane t holds d feat[8t + 2nt + e] of rows g (c0, c1) and g+8 (c2, c3); tile 4: columns 0..3 = inputs 32..35
    got = np.zeros((16, 36))
    for nt in range(4):
        for e in range(2):
            got[fr.G, 8 * fr.T + 2 * nt + e] = dx[nt][e]
            got[fr.G + 8, 8 * fr.T + 2 * nt + e] = dx[nt][2 + e]
    sel = fr.T < 2
    for e in range(2):
        got[fr.G[sel], 32 + 2 * fr.T[sel] + e] = dx[4][e][sel]
        got[fr.G[sel] + 8, 32 + 2 * fr.T[sel] + e] = dx[4][2 + e][sel]
    np.testing.assert_allclose(got, dx_ref, atol=1e-11)


def _gate_frag(gate, nt, e):
    row = fr.G + 8 * (e >> 1)
    col = 8 * nt + 2 * fr.T + (e & 1)
    return gate[row, col]


@pytest.mark.parametrize("k", [1, 10, 16])
def test_part_b_products(k):
    rng = np.random.default_rng(100 + k)
    w1, b1, w2, b2 = _weights(k, rng)
    cb, cols, S = fr.col_plan(k)
    W1F, W2B = fr.build_W1F(w1), fr.build_W2B(w2, k)
    x = rng.standard_normal((16, 36))
    Xt = np.zeros(16 * fr.SX)
    Xt.reshape(16, fr.SX)[:, :36] = x
    Xt.reshape(16, fr.SX)[:, 36:] = rng.standard_normal((16, 5))     # the tile's other columns must not leak into a flushed value
    D = np.zeros((16, S))
    for m in range(4):
        D[:, cb[m]:cb[m] + fr.out_count(m, k)] = rng.standard_normal((16, fr.out_count(m, k)))
    Dt = D.reshape(-1)
    for m in range(4):
        n = fr.out_count(m, k)
        Hfull = np.maximum(x @ w1[m].T + b1[m], 0.0)
        dHfull = (D[:, cb[m]:cb[m] + n] @ w2[m]) * (Hfull > 0)
        for hh in range(2):
            acc1 = [[np.zeros(32) for _ in range(4)] for _ in range(5)]
            nmt = (n + 15) // 16
            acc2 = [[[np.zeros(32) for _ in range(4)] for _ in range(2)] for _ in range(nmt)]
            for sub in range(2):
                hT = fr.partB_hidden(Xt, sub, W1F, b1, m, hh)
                hT = [np.maximum(v, 0.0) for v in hT]   # (the model returns the pre-activation; the kernel applies the relu)
                HT = _c_tile(hT)         # rows h = 16hh + r, columns anchors 8sub + c
                np.testing.assert_allclose(HT, Hfull[8 * sub:8 * sub + 8, 16 * hh:16 * hh + 16].T, atol=1e-12)
                dT = fr.partB_dhidden(Dt, S, sub, W2B, k, m, hh)
                dT = [np.where(h > 0, d, 0.0) for h, d in zip(hT, dT)]
                np.testing.assert_allclose(_c_tile(dT), dHfull[8 * sub:8 * sub + 8, 16 * hh:16 * hh + 16].T, atol=1e-11)
                acc1 = fr.partB_dW1(acc1, dT, Xt, sub)
                acc2 = fr.partB_dW2(acc2, hT, Dt, S, sub, k, m)
            dW1 = np.concatenate([_c_tile(t) for t in acc1], axis=1)[:, :36]      # rows h = 16hh + r, columns i
            np.testing.assert_allclose(dW1, (dHfull.T @ x)[16 * hh:16 * hh + 16], atol=1e-10)
            dW2_ref = D[:, cb[m]:cb[m] + n].T @ Hfull       # [o][h]
            for mt in range(nmt):
                tile = np.concatenate([_c_tile(acc2[mt][0]), _c_tile(acc2[mt][1])], axis=1)     # rows o = 16mt + r, columns h = 16hh + c
                rows = min(16, n - 16 * mt)
                np.testing.assert_allclose(tile[:rows], dW2_ref[16 * mt:16 * mt + rows, 16 * hh:16 * hh + 16], atol=1e-10)
