"""Lane-level model of the operand maps of the decode kernels (gsr_decode.cu): every index formula the kernels use to build
mma.m16n8k8 fragments from shared memory / registers, written once in numpy with one array element per lane, so that the
formulas can be checked against plain matrix products on the CPU (tests/test_decode_fragments.py).  The CUDA source uses the
same names (in1, in1b, W1F, W1B, W2F, W2B, the OUT / X tiles) and the same expressions.

m16n8k8 (row.col), lane = 4 g + t:
    A  a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)
    B  b0 (k = t, n = g)  b1 (k = t+4, n = g)
    C  c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
"""
import numpy as np

LANE = np.arange(32)
G = LANE >> 2
T = LANE & 3
HID = 32
NIN = 36
SX = 41          # X tile row stride: [0..31] feat, [32..34] view dir, [35] dist, [36..39] their gradients, [40] anchor id


def mma(c, a, b):
    """c[4][32] += A B with the fragment maps above (a[4][32], b[2][32])."""
    A = np.zeros((16, 8)); B = np.zeros((8, 8)); C = np.zeros((16, 8))
    A[G, T] = a[0]; A[G + 8, T] = a[1]; A[G, T + 4] = a[2]; A[G + 8, T + 4] = a[3]
    B[T, G] = b[0]; B[T + 4, G] = b[1]
    C[G, 2 * T] = c[0]; C[G, 2 * T + 1] = c[1]; C[G + 8, 2 * T] = c[2]; C[G + 8, 2 * T + 1] = c[3]
    C = C + A @ B
    return [C[G, 2 * T], C[G, 2 * T + 1], C[G + 8, 2 * T], C[G + 8, 2 * T + 1]]


# ---- column plan of the OUT tile -----------------------------------------------------------------------------------------
def out_count(m, k):
    return (k, k, 7 * k, 3 * k)[m]


def col_plan(k):
    """cb[m]: first column of MLP m (each MLP padded to a multiple of 16 columns); cols; S (row stride, cols + 4)."""
    cb, c = [], 0
    for m in range(4):
        cb.append(c)
        c += (out_count(m, k) + 15) // 16 * 16
    return cb, c, c + 4


# ---- input-slot maps -----------------------------------------------------------------------------------------------------
def in1(ks, s):
    """layer-1 contraction slot s (0..7) of k-step ks (0..4) -> input index, -1 = zero pad.  Lane t holds feat[8t..8t+7] as two
    float4: slot t of k-step ks is F0[ks], slot t+4 is F1[ks]; k-step 4 carries (ux, uy, uz, dist) in slots 0..3."""
    if ks < 4:
        return 8 * (s & 3) + 4 * (s >> 2) + ks
    return 32 + s if s < 4 else -1


def in1b(nt, c):
    """layer-1 backward output column c (0..7) of n-tile nt (0..4) -> input index: lane t ends up with d feat[8t..8t+7] in
    tiles 0..3 (two float4 stores) and tile 4 carries the view / distance gradients in columns 0..3."""
    if nt < 4:
        return 8 * (c >> 1) + 2 * nt + (c & 1)
    return 32 + c if c < 4 else -1


# ---- shared-memory weight copies (built by load_weights) ------------------------------------------------------------------
def build_W1F(w1):
    """W1F[m][ks][p][lane][4]: (b0, b1) of n-tiles 2p and 2p+1:  b0 = W1[m][8nt+g][in1(ks,t)], b1 = W1[m][8nt+g][in1(ks,t+4)]."""
    out = np.zeros((4, 5, 2, 32, 4))
    for m in range(4):
        for ks in range(5):
            for p in range(2):
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    for e in range(4):
                        nt, slot = 2 * p + (e >> 1), t + 4 * (e & 1)
                        i = in1(ks, slot)
                        out[m, ks, p, lane, e] = 0.0 if i < 0 else w1[m][8 * nt + g, i]
    return out


def build_W1B(w1):
    """W1B[m][ksp][nt][lane][4]: (b0, b1) of k-steps 2ksp and 2ksp+1:  b0 = W1[m][8ks+2t][in1b(nt,g)], b1 = W1[m][8ks+2t+1][in1b(nt,g)]."""
    out = np.zeros((4, 2, 5, 32, 4))
    for m in range(4):
        for ksp in range(2):
            for nt in range(5):
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    for e in range(4):
                        ks, h = 2 * ksp + (e >> 1), 0
                        h = 8 * ks + 2 * t + (e & 1)
                        i = in1b(nt, g)
                        out[m, ksp, nt, lane, e] = 0.0 if i < 0 else w1[m][h, i]
    return out


def tile8_base(k):
    """first 8-column tile of each MLP in the padded column plan"""
    cb, _, _ = col_plan(k)
    return [c // 8 for c in cb]


def build_W2F(w2, k):
    """W2F[tile8][ksp][lane][4]: (b0, b1) of k-steps 2ksp, 2ksp+1:  b0 = W2[m][o = 8nt+g][8ks+2t], b1 = W2[m][o][8ks+2t+1] (0 for o >= n_m)."""
    cb, cols, _ = col_plan(k)
    out = np.zeros((cols // 8, 2, 32, 4))
    for m in range(4):
        for nt in range((cb[m + 1] if m < 3 else cols) // 8 - cb[m] // 8):
            for ksp in range(2):
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    o = 8 * nt + g
                    for e in range(4):
                        h = 8 * (2 * ksp + (e >> 1)) + 2 * t + (e & 1)
                        out[cb[m] // 8 + nt, ksp, lane, e] = w2[m][o, h] if o < out_count(m, k) else 0.0
    return out


def build_W2B(w2, k):
    """W2B[tile8][p][lane][4]: (b0, b1) of n-tiles 2p, 2p+1 (h):  b0 = W2[m][o = 8ks+t][8nt+g], b1 = W2[m][8ks+t+4][8nt+g] (0 for o >= n_m)."""
    cb, cols, _ = col_plan(k)
    out = np.zeros((cols // 8, 2, 32, 4))
    for m in range(4):
        for ks in range((cb[m + 1] if m < 3 else cols) // 8 - cb[m] // 8):
            for p in range(2):
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    for e in range(4):
                        o = 8 * ks + t + 4 * (e & 1)
                        h = 8 * (2 * p + (e >> 1)) + g
                        out[cb[m] // 8 + ks, p, lane, e] = w2[m][o, h] if o < out_count(m, k) else 0.0
    return out


# ---- part A: one warp, 16 anchors as the M dimension -----------------------------------------------------------------------
def x_fragments(x):
    """x[16][36] -> A fragments xa[ks][4][32] as the gather builds them (lane holds rows g and g+8)."""
    xa = []
    for ks in range(5):
        fr = []
        for e in range(4):
            row = G + 8 * (e & 1)
            slot = T + 4 * (e >> 1)
            idx = np.array([in1(ks, int(s)) for s in slot])
            fr.append(np.where(idx >= 0, x[row, np.maximum(idx, 0)], 0.0))
        xa.append(fr)
    return xa


def layer1_forward(xa, W1F, b1, m):
    """-> h[nt][4][32]: C fragments of relu(x W1[m]^T + b1[m]) (rows = anchors, columns = hidden units 8nt + 2t, +1)."""
    h = []
    for nt in range(4):
        h.append([b1[m][8 * nt + 2 * T], b1[m][8 * nt + 2 * T + 1], b1[m][8 * nt + 2 * T], b1[m][8 * nt + 2 * T + 1]])
    for ks in range(5):
        for p in range(2):
            e = W1F[m, ks, p, LANE]          # one LDS.128 per lane
            h[2 * p] = mma(h[2 * p], xa[ks], [e[:, 0], e[:, 1]])
            h[2 * p + 1] = mma(h[2 * p + 1], xa[ks], [e[:, 2], e[:, 3]])
    return [[np.maximum(v, 0.0) for v in tile] for tile in h]


def c_as_a(c):
    """a C fragment (row g, columns 2t / 2t+1) as the A fragment of the next product, contraction slot t <-> column 2t,
    slot t+4 <-> column 2t+1"""
    return [c[0], c[2], c[1], c[3]]


def layer2_forward(h, W2F, b2pad, k, m, OUT, S):
    cb, cols, _ = col_plan(k)
    nt8 = (out_count(m, k) + 7) // 8
    for nt in range(nt8):
        col = cb[m] + 8 * nt + 2 * T
        acc = [b2pad[col], b2pad[col + 1], b2pad[col], b2pad[col + 1]]
        for ksp in range(2):
            e = W2F[cb[m] // 8 + nt, ksp, LANE]
            acc = mma(acc, c_as_a(h[2 * ksp]), [e[:, 0], e[:, 1]])
            acc = mma(acc, c_as_a(h[2 * ksp + 1]), [e[:, 2], e[:, 3]])
        OUT[G * S + col] = acc[0]; OUT[G * S + col + 1] = acc[1]
        OUT[(G + 8) * S + col] = acc[2]; OUT[(G + 8) * S + col + 1] = acc[3]


def layer2_backward(OUT, S, W2B, k, m):
    """-> dh[nt][4][32]: C fragments of dOUT[:, MLP m] W2[m] (rows = anchors, columns = hidden 8nt + 2t, +1), not yet gated."""
    cb, _, _ = col_plan(k)
    dh = [[np.zeros(32) for _ in range(4)] for _ in range(4)]
    for ks in range((out_count(m, k) + 7) // 8):
        c0 = cb[m] + 8 * ks
        a = [OUT[G * S + c0 + T], OUT[(G + 8) * S + c0 + T], OUT[G * S + c0 + T + 4], OUT[(G + 8) * S + c0 + T + 4]]
        for p in range(2):
            e = W2B[cb[m] // 8 + ks, p, LANE]
            dh[2 * p] = mma(dh[2 * p], a, [e[:, 0], e[:, 1]])
            dh[2 * p + 1] = mma(dh[2 * p + 1], a, [e[:, 2], e[:, 3]])
    return dh


def layer1_backward(dx, dh, W1B, m):
    """dx[nt][4][32] += dh W1[m]; afterwards lane t holds d feat[8t + 2nt + e] of rows g (c0, c1) and g+8 (c2, c3)."""
    for ksp in range(2):
        for nt in range(5):
            e = W1B[m, ksp, nt, LANE]
            dx[nt] = mma(dx[nt], c_as_a(dh[2 * ksp]), [e[:, 0], e[:, 1]])
            dx[nt] = mma(dx[nt], c_as_a(dh[2 * ksp + 1]), [e[:, 2], e[:, 3]])
    return dx


# ---- part B: warp (m, hh) over n-tiles of 8 anchors, hidden units 16hh .. 16hh+15 as the M dimension ---------------------------
def partB_hidden(Xt, sub, W1F, b1, m, hh):
    """-> hT[4][32]: C fragment of relu(W1[m] x^T + b1) for rows h = 16hh + g (+8), columns = anchors 8sub + 2t (+1)."""
    acc = [b1[m][16 * hh + G], b1[m][16 * hh + G], b1[m][16 * hh + G + 8], b1[m][16 * hh + G + 8]]
    for ks in range(5):
        e = W1F[m, ks, hh, LANE]
        a = [e[:, 0], e[:, 2], e[:, 1], e[:, 3]]
        i0 = np.array([in1(ks, int(s)) for s in T]); i1 = np.array([in1(ks, int(s) + 4) for s in T])
        row = (8 * sub + G) * SX
        b0 = Xt[row + i0]
        b1f = np.where(i1 >= 0, Xt[row + np.maximum(i1, 0)], 0.0)
        acc = mma(acc, a, [b0, b1f])
    return acc


def partB_dhidden(Dt, S, sub, W2B, k, m, hh):
    """-> dT[4][32]: C fragment of W2[m]^T dOUT^T, rows h = 16hh + g (+8), columns = anchors 8sub + 2t (+1); not gated."""
    cb, _, _ = col_plan(k)
    acc = [np.zeros(32) for _ in range(4)]
    for ks in range((out_count(m, k) + 7) // 8):
        e = W2B[cb[m] // 8 + ks, hh, LANE]
        a = [e[:, 0], e[:, 2], e[:, 1], e[:, 3]]
        base = (8 * sub + G) * S + cb[m] + 8 * ks
        acc = mma(acc, a, [Dt[base + T], Dt[base + T + 4]])
    return acc


def partB_dW1(acc1, dT, Xt, sub):
    """acc1[nt][4][32] += dT (as A: rows h, contraction = the 8 anchors) X; rows h = 16hh + g (+8), columns i = 8nt + 2t (+1)."""
    a = c_as_a(dT)
    for nt in range(5):
        col = 8 * nt + G                      # columns 36..39 of tile 4 are not inputs: their accumulators are never flushed
        b0 = Xt[(8 * sub + 2 * T) * SX + col]
        b1 = Xt[(8 * sub + 2 * T + 1) * SX + col]
        acc1[nt] = mma(acc1[nt], a, [b0, b1])
    return acc1


def partB_dW2(acc2, hT, Dt, S, sub, k, m):
    """acc2[mt][nt][4][32] += dOUT^T (A: rows o = 16mt + g (+8), contraction = anchors) H (B from the hT fragment);
    columns h = 16hh + 8nt + 2t (+1)."""
    cb, _, _ = col_plan(k)
    for mt in range((out_count(m, k) + 15) // 16):
        c0 = cb[m] + 16 * mt + G
        r0 = (8 * sub + 2 * T) * S
        r1 = (8 * sub + 2 * T + 1) * S
        a = [Dt[r0 + c0], Dt[r0 + c0 + 8], Dt[r1 + c0], Dt[r1 + c0 + 8]]
        acc2[mt][0] = mma(acc2[mt][0], a, [hT[0], hT[1]])
        acc2[mt][1] = mma(acc2[mt][1], a, [hT[2], hT[3]])
    return acc2
