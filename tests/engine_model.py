"""Numpy model of the CUDA engine's ALGORITHM (not of the reference's): the same chunked
affine-scan Crank-Nicolson, the same pair-local kernel schedule and the same cross-step fusion of
the even rotations that ionization_b200/csrc/engine.cu implements.  It exists so the algebra can be
checked on a machine without a GPU (tests/test_engine_model.py compares it with the golden
fixtures); it is test code, never imported by the product.
"""
import numpy as np


def factor(h_diag, h_off, tau, M, T):
    """per channel: w_i = 1/p_i, e_i = -i tau off_i w_i, chunk aggregates P_t, Q_t (engine.cu: k_factor)."""
    L, R = h_diag.shape
    Rp = M * T
    toff = np.zeros(Rp)
    toff[: R - 1] = tau * h_off
    w = np.ones((L, Rp), dtype=np.complex128)
    D = 1 + 1j * tau * h_diag
    for l in range(L):
        p = D[l, 0]
        w[l, 0] = 1 / p
        for i in range(1, R):
            O = 1j * toff[i - 1]
            p = D[l, i] - O * O * w[l, i - 1]
            w[l, i] = 1 / p
    e = -1j * toff[None, :] * w
    e_prev = np.concatenate([np.zeros((L, 1)), e[:, :-1]], axis=1)  # e_{i-1}
    P = np.prod(e_prev.reshape(L, T, M), axis=2)  # prod_{i=tM-1}^{tM+M-2} e_i
    Q = np.prod(e.reshape(L, T, M), axis=2)
    return w, e, P, Q


def _scan_exclusive_fwd(P, B):
    """Y_in[t] for maps f_t(v) = P_t v + B_t, Y_{-1} = 0 (Kogge-Stone in the engine; serial here)."""
    Yin = np.zeros_like(B)
    acc = 0
    for t in range(len(B)):
        Yin[t] = acc
        acc = P[t] * acc + B[t]
    return Yin


def cn_channel(g, w, e, P, Q, M, T):
    """x = 2 (1 + i tau H)^-1 g - g on one padded channel (length M*T), chunk t = rows [tM, tM+M)."""
    G = g.reshape(T, M)
    W = w.reshape(T, M)
    E = e.reshape(T, M)
    elink = np.concatenate([[0], e[M - 1 :: M][:-1]])  # e_{tM-1}
    # forward pass 1 (y_in = 0)
    z = G[:, 0].copy()
    for k in range(1, M):
        z = G[:, k] + E[:, k - 1] * z
    Yin = _scan_exclusive_fwd(P, z)
    # forward pass 2
    U = np.zeros_like(G)
    y = G[:, 0] + elink * Yin
    U[:, 0] = W[:, 0] * y
    for k in range(1, M):
        y = G[:, k] + E[:, k - 1] * y
        U[:, k] = W[:, k] * y
    # backward pass 1
    z = U[:, M - 1].copy()
    for k in range(M - 2, -1, -1):
        z = U[:, k] + E[:, k] * z
    Xin = _scan_exclusive_fwd(Q[::-1], z[::-1])[::-1]
    out = np.zeros_like(G)
    x = U[:, M - 1] + E[:, M - 1] * Xin
    out[:, M - 1] = 2 * x - G[:, M - 1]
    for k in range(M - 2, -1, -1):
        x = U[:, k] + E[:, k] * x
        out[:, k] = 2 * x - G[:, k]
    return out.reshape(-1)


def _pad(a, Rp):
    out = np.zeros(a.shape[:-1] + (Rp,), dtype=a.dtype)
    out[..., : a.shape[-1]] = a
    return out


def rot_pairs(g, parity, ang, real):
    """pair-local rotation of channel pairs (l, l+1), l % 2 == parity; ang: (L-1, Rp)."""
    L = g.shape[0]
    for l in range(parity, L - 1, 2):
        c, s = np.cos(ang[l]), np.sin(ang[l])
        a, b = g[l].copy(), g[l + 1].copy()
        if real:
            g[l], g[l + 1] = c * a + s * b, -s * a + c * b
        else:
            g[l], g[l + 1] = c * a - 1j * s * b, -1j * s * a + c * b


def h2_pairs(g, l_parity, order, th, R):
    """Hadamard + the two r-sublayers + Hadamard back on l-pairs of parity l_parity.
    order = (first r parity, second r parity); th: (L-1, Rp) zero where no r-pair starts."""
    L = g.shape[0]
    rs2 = 1 / np.sqrt(2)
    for l in range(l_parity, L - 1, 2):
        s = (g[l] + g[l + 1]) * rs2
        d = (g[l] - g[l + 1]) * rs2
        for rp in order:
            jj = np.arange(rp, R - 1, 2)
            c, sn = np.cos(th[l, jj]), np.sin(th[l, jj])
            s0, s1, d0, d1 = s[jj].copy(), s[jj + 1].copy(), d[jj].copy(), d[jj + 1].copy()
            s[jj], s[jj + 1] = c * s0 + sn * s1, -sn * s0 + c * s1
            d[jj], d[jj + 1] = c * d0 - sn * d1, sn * d0 + c * d1
        g[l], g[l + 1] = (s + d) * rs2, (s - d) * rs2


def run_sh_model(p, M=4, fused=True):
    """engine schedule for L even (fast path).  Returns final g (L, R)."""
    kind = str(p["kind"])
    L, R = int(p["L"]), int(p["R"])
    assert L % 2 == 0
    T = -(-R // M)
    T = -(-T // 32) * 32
    Rp = M * T
    g = _pad(np.array(p["g0"], dtype=np.complex128), Rp)
    mask = _pad(np.asarray(p["mask"], dtype=np.float64), Rp)
    taus, fields = p["taus"], p["fields"]
    tau0 = taus[0]
    w, e, P, Q = factor(np.asarray(p["h_diag"]), np.asarray(p["h_off"]), tau0, M, T)
    c_l = np.asarray(p["c_l"])
    N = len(taus)
    s = taus * fields

    def cn_all(g):
        for l in range(L):
            g[l] = cn_channel(g[l], w[l], e[l], P[l], Q[l], M, T)

    if kind == "sh_len_so":
        x = _pad(np.asarray(p["x_j"]), Rp)
        base = c_l[:, None] * x[None, :]
        rot_pairs(g, 0, s[0] * base, False)
        for n in range(N):
            rot_pairs(g, 1, s[n] * base, False)
            cn_all(g)
            rot_pairs(g, 1, s[n] * base, False)
            if fused and n + 1 < N:
                rot_pairs(g, 0, (s[n] + s[n + 1]) * base, False)
                g *= mask[None, :]
            else:
                rot_pairs(g, 0, s[n] * base, False)
                g *= mask[None, :]
                if n + 1 < N:
                    rot_pairs(g, 0, s[n + 1] * base, False)
    elif kind == "sh_vel_so":
        y = _pad(np.asarray(p["y_j"]), Rp)
        z = _pad(np.asarray(p["z_j"]), Rp)
        b1 = np.asarray(p["f1_l"])[:, None] * y[None, :]
        b2 = c_l[:, None] * z[None, :]
        rot_pairs(g, 0, s[0] * b1, True)
        for n in range(N):
            rot_pairs(g, 1, s[n] * b1, True)  # C
            h2_pairs(g, 0, (0, 1), s[n] * b2, R)  # B': ee, eo
            h2_pairs(g, 1, (0, 1), s[n] * b2, R)  # A: oe, oo
            cn_all(g)
            h2_pairs(g, 1, (1, 0), s[n] * b2, R)  # A: oo, oe
            h2_pairs(g, 0, (1, 0), s[n] * b2, R)  # B: eo, ee
            rot_pairs(g, 1, s[n] * b1, True)  # C
            if fused and n + 1 < N:
                rot_pairs(g, 0, (s[n] + s[n + 1]) * b1, True)  # D
                g *= mask[None, :]
            else:
                rot_pairs(g, 0, s[n] * b1, True)
                g *= mask[None, :]
                if n + 1 < N:
                    rot_pairs(g, 0, s[n + 1] * b1, True)
    else:
        raise ValueError(kind)
    return g[:, :R]


# ---------------------------------------------------------------------------------------------
# ION_SH_LEN_ADI (csrc/adi.cuh): per radial position an L x L tridiagonal solve along l, cut into chunks of CL channels
# ---------------------------------------------------------------------------------------------
def adi_l_pass(g1, beta, CL=8):
    """g1: (L,) right-hand side at one radial position; beta: (L-1,) real couplings (tau E c_l x_j).
    Returns (1 - i B)(1 + i B)^-1 g1 the way k_adi_l computes it: real pivots through composed Moebius maps,
    chunked affine recurrences with sequential prefixes over the chunk aggregates, 2x - g1 at the end."""
    L = len(g1)
    NC = -(-L // CL)
    bq = np.zeros(NC * CL + 1)  # bq[l]: coupling of channel l with l-1
    bq[1:L] = beta
    # chunk matrices of w -> 1 / (1 + a w)
    mats = []
    for c in range(NC):
        A, B, C, D = 1.0, 0.0, 0.0, 1.0
        for l in range(c * CL, min(L, (c + 1) * CL)):
            a = bq[l] ** 2
            A, B, C, D = C, D, a * A + C, a * B + D
        mats.append((A, B, C, D))
    w = np.ones(NC * CL)
    w_in = np.ones(NC)
    for c in range(NC):
        wi = 1.0
        for kk in range(c):
            A, B, C, D = mats[kk]
            wi = (A * wi + B) / (C * wi + D)
        w_in[c] = wi
        wp = wi
        for l in range(c * CL, min(L, (c + 1) * CL)):
            wp = 1.0 / (bq[l] ** 2 * wp + 1.0)
            w[l] = wp
    gp = np.zeros(NC * CL, dtype=np.complex128)
    gp[:L] = g1
    ef = np.zeros(NC * CL)
    for c in range(NC):
        ef[c * CL] = bq[c * CL] * w_in[c]
        for k in range(1, CL):
            ef[c * CL + k] = bq[c * CL + k] * w[c * CL + k - 1]
    # forward
    Mc, Yc = [], []
    for c in range(NC):
        z, m = gp[c * CL], -1j * ef[c * CL]
        for k in range(1, CL):
            z = gp[c * CL + k] - 1j * ef[c * CL + k] * z
            m = -1j * ef[c * CL + k] * m
        Mc.append(m)
        Yc.append(z)
    y = np.zeros(NC * CL, dtype=np.complex128)
    for c in range(NC):
        yin = 0
        for kk in range(c):
            yin = Mc[kk] * yin + Yc[kk]
        prev = yin
        for k in range(CL):
            prev = gp[c * CL + k] - 1j * ef[c * CL + k] * prev
            y[c * CL + k] = prev
    # backward
    eb = bq[1:] * w
    Qc, Xc = [], []
    for c in range(NC):
        hi = c * CL + CL - 1
        z, m = w[hi] * y[hi], -1j * eb[hi]
        for l in range(hi - 1, c * CL - 1, -1):
            z = w[l] * y[l] - 1j * eb[l] * z
            m = -1j * eb[l] * m
        Qc.append(m)
        Xc.append(z)
    out = np.zeros(NC * CL, dtype=np.complex128)
    for c in range(NC):
        xin = 0
        for kk in range(NC - 1, c, -1):
            xin = Qc[kk] * xin + Xc[kk]
        x = xin
        for l in range(c * CL + CL - 1, c * CL - 1, -1):
            x = w[l] * y[l] - 1j * eb[l] * x
            out[l] = 2 * x - gp[l]
    return out[:L]


def run_sh_adi_model(p, CL=8):
    """engine schedule of ION_SH_LEN_ADI: [explicit r + l pass] then [(1 + i tau H0)^-1 = (CN + 1)/2, mask]."""
    L, R = int(p["L"]), int(p["R"])
    g = np.array(p["g0"], dtype=np.complex128)
    hd, ho = np.asarray(p["h_diag"]), np.asarray(p["h_off"])
    c_l, x_j, mask = np.asarray(p["c_l"]), np.asarray(p["x_j"]), np.asarray(p["mask"])
    M = 4
    T = -(-(-(-R // M)) // 32) * 32
    for tau, E in zip(p["taus"], p["fields"]):
        g1 = (1 - 1j * tau * hd) * g
        g1[:, 1:] += (-1j * tau * ho) * g[:, :-1]
        g1[:, :-1] += (-1j * tau * ho) * g[:, 1:]
        g2 = np.empty_like(g1)
        for j in range(R):
            g2[:, j] = adi_l_pass(g1[:, j], tau * E * c_l * x_j[j], CL)
        w, e, P, Q = factor(hd, ho, tau, M, T)
        for l in range(L):
            row = _pad(g2[l][None, :], M * T)[0]
            g2[l] = (0.5 * (cn_channel(row, w[l], e[l], P[l], Q[l], M, T) + row))[:R]
        g = g2 * mask[None, :]
    return g
