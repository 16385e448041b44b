"""ORACLE TEST INFRASTRUCTURE -- not product code.

CPU restatement (numpy) of the reference's mesh time-evolution hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker.

PARITY PINNING: every function here is checked in ``tests/test_oracle_pinned.py``
against (a) fixtures produced by the UNMODIFIED reference run in this container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``), (b) the reference's own
known answers -- the six final overlaps in
``dev/meshes/mesh_refactoring_helper.py:204-251`` -- and (c) the compiled
``cy.tdma`` (``oracle/_ref``) when present.

All file:line citations are relative to /root/reference/ionization.

Conventions shared with the CUDA engine (see DESIGN.md "data layout"):
  * ``g`` for SphericalHarmonicMesh has shape ``(L, R)``, C-contiguous, r
    fastest -- the reference's own storage (mesh/meshes.py:1003, :1052-1064).
  * The field-free Hamiltonian is tridiagonal in r per l-channel:
    ``h_diag[l, j]`` and ``h_off[j]`` (coupling j<->j+1, identical for all l;
    mesh/mesh_operators.py:889-928 zeroes it across l-block edges).
  * per step ``tau = dt / (2 hbar)`` (mesh/evolution_methods.py:92).
"""
import numpy as np


# --------------------------------------------------------------------------
# a1: Thomas algorithm                                         cy.pyx:9-50
# --------------------------------------------------------------------------
def tdma(sub, diag, sup, d):
    """x = M^-1 d for tridiagonal M, no pivoting.  Follows cy.pyx:28-48.

    ``sub[i]`` multiplies x[i-1] in row i (i>=1), ``sup[i]`` multiplies x[i+1]
    in row i (i<=n-2); both have length n-1.
    """
    n = len(d)
    sub = np.asarray(sub, dtype=np.complex128)
    diag = np.asarray(diag, dtype=np.complex128)
    sup = np.asarray(sup, dtype=np.complex128)
    d = np.asarray(d, dtype=np.complex128)
    new_sup = np.zeros(max(n - 1, 0), dtype=np.complex128)
    new_d = np.zeros(n, dtype=np.complex128)
    x = np.zeros(n, dtype=np.complex128)
    if n == 1:
        x[0] = d[0] / diag[0]
        return x
    new_sup[0] = sup[0] / diag[0]  # cy.pyx:29
    new_d[0] = d[0] / diag[0]  # cy.pyx:30
    for i in range(1, n - 1):  # cy.pyx:31-38
        denom = diag[i] - sub[i - 1] * new_sup[i - 1]
        new_sup[i] = sup[i] / denom
        new_d[i] = (d[i] - sub[i - 1] * new_d[i - 1]) / denom
    new_d[n - 1] = (d[n - 1] - sub[n - 2] * new_d[n - 2]) / (diag[n - 1] - sub[n - 2] * new_sup[n - 2])  # :40
    x[n - 1] = new_d[n - 1]  # cy.pyx:43
    for i in range(n - 2, -1, -1):  # cy.pyx:44-45
        x[i] = new_d[i] - new_sup[i] * x[i + 1]
    return x


def tdma_batched(sub, diag, sup, d):
    """Vectorised-over-systems Thomas: arrays of shape (batch, n-1|n).  Same
    recurrence as :func:`tdma`, row loop in python, systems in numpy."""
    d = np.asarray(d, dtype=np.complex128)
    B, n = d.shape
    sub = np.broadcast_to(np.asarray(sub, dtype=np.complex128), (B, n - 1))
    sup = np.broadcast_to(np.asarray(sup, dtype=np.complex128), (B, n - 1))
    diag = np.broadcast_to(np.asarray(diag, dtype=np.complex128), (B, n))
    cp = np.zeros((B, max(n - 1, 1)), dtype=np.complex128)
    dp = np.zeros((B, n), dtype=np.complex128)
    x = np.zeros((B, n), dtype=np.complex128)
    if n == 1:
        return d / diag
    cp[:, 0] = sup[:, 0] / diag[:, 0]
    dp[:, 0] = d[:, 0] / diag[:, 0]
    for i in range(1, n - 1):
        denom = diag[:, i] - sub[:, i - 1] * cp[:, i - 1]
        cp[:, i] = sup[:, i] / denom
        dp[:, i] = (d[:, i] - sub[:, i - 1] * dp[:, i - 1]) / denom
    dp[:, n - 1] = (d[:, n - 1] - sub[:, n - 2] * dp[:, n - 2]) / (diag[:, n - 1] - sub[:, n - 2] * cp[:, n - 2])
    x[:, n - 1] = dp[:, n - 1]
    for i in range(n - 2, -1, -1):
        x[:, i] = dp[:, i] - cp[:, i] * x[:, i + 1]
    return x


# --------------------------------------------------------------------------
# a6: SphericalHarmonic field-free Hamiltonian    mesh_operators.py:841-928, :244-269
# --------------------------------------------------------------------------
def sh_alpha(j):
    """mesh_operators.py:841-843"""
    j = np.asarray(j, dtype=np.float64)
    x = j ** 2 + 2 * j
    return (x + 1) / (x + 0.75)


def sh_beta(j):
    """mesh_operators.py:845-847"""
    j = np.asarray(j, dtype=np.float64)
    x = 2 * j ** 2 + 2 * j
    return (x + 1) / (x + 0.5)


def sh_c_l(l):
    """mesh_operators.py:853-855"""
    l = np.asarray(l, dtype=np.float64)
    return (l + 1) / np.sqrt((2 * l + 1) * (2 * l + 3))


def sh_r_grid(r_bound, r_points):
    """mesh/meshes.py:1009-1011: r = linspace(0, r_bound, R) + delta_r/2."""
    r = np.linspace(0, r_bound, r_points)
    delta_r = r[1] - r[0]
    return r + delta_r / 2, delta_r


def sh_h0(r, delta_r, l_bound, potential_r, *, hbar, mass_kinetic, bohr_radius, l0_correction=True):
    """(h_diag[L,R], h_off[R-1]) of the LAGRANGIAN-derived radial Hamiltonian.

    Follows kinetic_energy_from_lagrangian (mesh_operators.py:889-928) plus the
    potential added on the diagonal by internal_hamiltonian (:244-269).
    ``mass_kinetic`` is electron_mass_reduced in the reference (:892-894).
    """
    R = len(r)
    j = np.arange(R)
    pre = -(hbar ** 2) / (2 * mass_kinetic * delta_r ** 2)  # :892-894
    beta = np.tile(sh_beta(j), (l_bound, 1)).astype(np.complex128)  # :898-900
    if l0_correction:
        dr = delta_r / bohr_radius
        beta[0, 0] += dr * (1 + dr) / 8  # :901-903 (flat index 0 only)
    h_diag = beta * (-2 * pre)  # :909
    h_off = sh_alpha(j[:-1]) * pre  # :905-910
    l = np.arange(l_bound)
    h_diag = h_diag + ((hbar ** 2) / (2 * mass_kinetic)) * (l * (l + 1))[:, None] / (r[None, :] ** 2)  # :912-920
    h_diag = h_diag + np.asarray(potential_r)[None, :]  # :244-269
    return h_diag, h_off.astype(np.float64)


# --------------------------------------------------------------------------
# a10: radial cosine mask                              potentials/masks.py:76-89
# --------------------------------------------------------------------------
def radial_cosine_mask(r, inner_radius, outer_radius, smoothness=8):
    r = np.asarray(r, dtype=np.float64)
    ramp = np.abs(np.cos(0.5 * np.pi * (r - inner_radius) / np.abs(outer_radius - inner_radius))) ** (1 / smoothness)
    return np.where((r >= inner_radius) & (r < outer_radius), ramp, np.where(r >= outer_radius, 0.0, 1.0))


# --------------------------------------------------------------------------
# Crank-Nicolson in r                     evolution_methods.py:98-111 + cy.tdma
# --------------------------------------------------------------------------
def cn_r(g, h_diag, h_off, tau):
    """g <- (1 + i tau H0)^-1 (1 - i tau H0) g, independently per channel.

    explicit half = DotOperator (mesh_operators.py:104-106), implicit half =
    TDMAOperator -> cy.tdma (:109-111).  ``g``: (L,R) or (R,).
    """
    g2 = np.atleast_2d(g)
    hd = np.atleast_2d(h_diag)
    off = np.asarray(h_off)
    rhs = (1 - 1j * tau * hd) * g2
    rhs[:, 1:] += (-1j * tau * off) * g2[:, :-1]
    rhs[:, :-1] += (-1j * tau * off) * g2[:, 1:]
    x = tdma_batched(1j * tau * off, 1 + 1j * tau * hd, 1j * tau * off, rhs)
    return x.reshape(np.shape(g))


# --------------------------------------------------------------------------
# a7: length-gauge l<->l+1 rotations          mesh_operators.py:988-1080
# --------------------------------------------------------------------------
def _flat_parity(L, R):
    """parity of the F-order flat index k = j*L + l of the LOWER member of the
    pair (l, l+1) at radius j -- mesh_operators.py:1045 (a[::2], a[1::2]) on an
    L-wrapped vector (mesh/meshes.py:1052-1064).  Shape (L-1, R)."""
    return (np.arange(R)[None, :] * L + np.arange(L - 1)[:, None]) % 2


def _rot_pairs(g, m00, m01, m10, m11, sel):
    """apply [[m00,m01],[m10,m11]] to (g[l], g[l+1]) where sel[l,j]."""
    out = g.copy()
    lo, hi = g[:-1], g[1:]
    nlo = m00 * lo + m01 * hi
    nhi = m10 * lo + m11 * hi
    out[:-1][sel] = nlo[sel]
    out[1:][sel] = nhi[sel]
    return out


def sh_len_sweep(g, angle, parity):
    """one DotOperator of split_interaction_operators (mesh_operators.py:1037-1080):
    [[cos a, -i sin a], [-i sin a, cos a]] on the pairs of the given flat parity.
    ``angle``: (L-1, R)."""
    L, R = g.shape
    c, s = np.cos(angle), np.sin(angle)
    return _rot_pairs(g, c, -1j * s, -1j * s, c, _flat_parity(L, R) == parity)


def sh_len_angle(tau, efield, c_l, x_j):
    """a = tau * E * (-q) * r_j * c_l (mesh_operators.py:988-1018, :1043).
    ``x_j = -q * r_j``; ``c_l``: length L-1."""
    return (tau * efield) * c_l[:, None] * x_j[None, :]


def sh_len_so_step(g, h_diag, h_off, c_l, x_j, mask, tau, efield):
    """One SplitInteractionOperator step, length gauge (evolution_methods.py:89-123)
    followed by the mask (mesh/meshes.py:251-257).  efield = E(t_{n+1} + dt/2)."""
    a = sh_len_angle(tau, efield, c_l, x_j)
    g = sh_len_sweep(g, a, 0)
    g = sh_len_sweep(g, a, 1)
    g = cn_r(g, h_diag, h_off, tau)
    g = sh_len_sweep(g, a, 1)
    g = sh_len_sweep(g, a, 0)
    return g * mask[None, :]


# --------------------------------------------------------------------------
# a8: velocity-gauge sweeps                    mesh_operators.py:1143-1408, :150-204
# --------------------------------------------------------------------------
def sh_vel_h1_sweep(g, theta1, parity):
    """split_h1 (mesh_operators.py:1204-1245): real rotation
    [[cos, sin], [-sin, cos]] on flat-parity l-pairs.  theta1: (L-1, R)."""
    L, R = g.shape
    c, s = np.cos(theta1), np.sin(theta1)
    return _rot_pairs(g, c, s, -s, c, _flat_parity(L, R) == parity)


def sh_vel_h2_sweep(g, theta2, l_parity, r_parity):
    """one SimilarityOperator of split_h2 (mesh_operators.py:1247-1408, :150-204).

    Hadamard over l-pairs of parity ``l_parity`` (by l), then on r-pairs of
    parity ``r_parity`` (by j, inside each block) rotate the sum row by +theta2
    and the difference row by -theta2, Hadamard back.  theta2: (L-1, R-1).
    Rows outside any pair (l=0 for odd sweeps, last row when unpaired) are
    multiplied by sqrt2/sqrt2 = 1 and their r-rotation is the identity.
    """
    L, R = g.shape
    out = g.copy()
    rs2 = 1 / np.sqrt(2)
    for l in range(l_parity, L - 1, 2):
        s = (g[l] + g[l + 1]) * rs2
        d = (g[l] - g[l + 1]) * rs2
        s2, d2 = s.copy(), d.copy()
        jj = np.arange(r_parity, R - 1, 2)
        c, sn = np.cos(theta2[l, jj]), np.sin(theta2[l, jj])
        s2[jj] = c * s[jj] + sn * s[jj + 1]
        s2[jj + 1] = -sn * s[jj] + c * s[jj + 1]
        d2[jj] = c * d[jj] - sn * d[jj + 1]
        d2[jj + 1] = sn * d[jj] + c * d[jj + 1]
        out[l] = (s2 + d2) * rs2
        out[l + 1] = (s2 - d2) * rs2
    return out


def sh_vel_angles(tau, vamp, f1_l, y_j, c_l, z_j):
    """theta1 = tau*hbar*(q/m)*A*c_l*(l+1)/r_j  (mesh_operators.py:1147-1161, :1209)
    theta2 = tau*hbar*(q/m)*A/(2 dr)*c_l*alpha_j (:1163-1178, :1252).
    ``f1_l = c_l*(l+1)``, ``y_j = hbar*(q/m)/r_j``, ``z_j = hbar*(q/m)/(2 dr)*alpha_j``."""
    s = tau * vamp
    return s * f1_l[:, None] * y_j[None, :], s * c_l[:, None] * z_j[None, :]


def sh_vel_so_step(g, h_diag, h_off, c_l, f1_l, y_j, z_j, mask, tau, vamp):
    """One SplitInteractionOperator step, velocity gauge.  Order
    (mesh_operators.py:1190-1202, evolution_methods.py:117-121):
    h1_e, h1_o, h2_ee, h2_eo, h2_oe, h2_oo, CN, reversed.  vamp = A(t_0..t_{n+1})."""
    th1, th2 = sh_vel_angles(tau, vamp, f1_l, y_j, c_l, z_j)
    fwd = [
        lambda g: sh_vel_h1_sweep(g, th1, 0),
        lambda g: sh_vel_h1_sweep(g, th1, 1),
        lambda g: sh_vel_h2_sweep(g, th2, 0, 0),
        lambda g: sh_vel_h2_sweep(g, th2, 0, 1),
        lambda g: sh_vel_h2_sweep(g, th2, 1, 0),
        lambda g: sh_vel_h2_sweep(g, th2, 1, 1),
    ]
    for f in fwd:
        g = f(g)
    g = cn_r(g, h_diag, h_off, tau)
    for f in reversed(fwd):
        g = f(g)
    return g * mask[None, :]


# --------------------------------------------------------------------------
# a5: SphericalHarmonic ADI, length gauge ("next" row f-4)
#     evolution_methods.py:49-77, mesh_operators.py:1020-1035
# --------------------------------------------------------------------------
def sh_len_adi_step(g, h_diag, h_off, c_l, x_j, mask, tau, efield):
    """[(1-i tau H0)_R, (1+i tau Hint)^-1_L, (1-i tau Hint)_L, (1+i tau H0)^-1_R], mask.
    Hint couples (l,l+1) at radius j with E*c_l*x_j, tridiagonal in l, zero diagonal."""
    L, R = g.shape
    hd = np.asarray(h_diag)
    # explicit in r
    rhs = (1 - 1j * tau * hd) * g
    rhs[:, 1:] += (-1j * tau * h_off) * g[:, :-1]
    rhs[:, :-1] += (-1j * tau * h_off) * g[:, 1:]
    # implicit then explicit in l: systems over l for each j
    w = efield * c_l[:, None] * x_j[None, :]  # (L-1, R)
    gl = rhs.T  # (R, L)
    wl = w.T  # (R, L-1)
    gl = tdma_batched(1j * tau * wl, np.ones((R, L), dtype=np.complex128), 1j * tau * wl, gl)
    ex = gl.copy()
    ex[:, 1:] += (-1j * tau * wl) * gl[:, :-1]
    ex[:, :-1] += (-1j * tau * wl) * gl[:, 1:]
    g = ex.T
    # implicit in r
    g = tdma_batched(1j * tau * h_off, 1 + 1j * tau * hd, 1j * tau * h_off, g)
    return g * mask[None, :]


# --------------------------------------------------------------------------
# a9: LineMesh                         mesh_operators.py:304-427, evolution_methods.py:49-77
# --------------------------------------------------------------------------
def line_h0(z, potential_z, *, hbar, mass):
    """H0 = -hbar^2/(2 m dz^2) tridiag(1,-2,1) + diag(V)  (mesh_operators.py:310-318, :244-269)."""
    dz = np.abs(z[1] - z[0])
    pre = -(hbar ** 2) / (2 * mass * dz ** 2)
    h_diag = (-2 * pre) * np.ones(len(z), dtype=np.complex128) + np.asarray(potential_z)
    h_off = pre * np.ones(len(z) - 1)
    return h_diag, h_off


def line_cn_len_step(g, h_diag, h_off, w_z, mask, tau, efield):
    """AlternatingDirectionImplicit on LineMesh, length gauge (evolution_methods.py:49-77,
    mesh_operators.py:271-298, :320-327): H = H0 + diag(-q z E(t_{n+1})); w_z = -q*z."""
    return cn_r(g, h_diag + efield * w_z, h_off, tau) * mask


def line_so_len_step(g, h_diag, h_off, w_z, mask, tau, efield):
    """SplitInteractionOperator on LineMesh, length gauge (mesh_operators.py:329-341):
    P CN(H0) P with P = exp(-i tau (-q z E))."""
    p = np.exp(-1j * tau * efield * w_z)
    return p * cn_r(p * g, h_diag, h_off, tau) * mask


def line_vel_sweep(g, theta, parity):
    """mesh_operators.py:384-427: [[cos, sin], [-sin, cos]] on (z_k, z_k+1), k of given parity."""
    out = g.copy()
    k = np.arange(parity, len(g) - 1, 2)
    c, s = np.cos(theta), np.sin(theta)
    out[k] = c * g[k] + s * g[k + 1]
    out[k + 1] = -s * g[k] + c * g[k + 1]
    return out


def line_so_vel_step(g, h_diag, h_off, v_pref, mask, tau, vamp):
    """SplitInteractionOperator on LineMesh, velocity gauge: theta = tau*hbar*(q/m)*A/(2 dz)
    (mesh_operators.py:358-427); v_pref = hbar*(q/m)/(2 dz)."""
    theta = tau * vamp * v_pref
    g = line_vel_sweep(g, theta, 0)
    g = line_vel_sweep(g, theta, 1)
    g = cn_r(g, h_diag, h_off, tau)
    g = line_vel_sweep(g, theta, 1)
    g = line_vel_sweep(g, theta, 0)
    return g * mask


# --------------------------------------------------------------------------
# a11: observables                       mesh/meshes.py:195-237, :1099-1136
# --------------------------------------------------------------------------
def norm(g, ipm):
    """mesh/meshes.py:195-217 with the empty operator sum."""
    return float(np.real(np.sum(np.conj(g) * g) * ipm))


def inner_product_rows(g, state_l, state_rows, ipm):
    """SphericalHarmonicMesh.inner_product shortcut (mesh/meshes.py:1117-1129):
    sum_j conj(R_s[j]) g[s.l, j] * delta_r for each single-l test state."""
    return np.array([np.sum(np.conj(row) * g[l]) * ipm for l, row in zip(state_l, state_rows)])


def inner_product_full(a, g, ipm):
    """QuantumMesh.inner_product (mesh/meshes.py:195-200)."""
    return np.sum(np.conj(a) * g) * ipm


def norm_by_l(g, ipm):
    """mesh/meshes.py:1133-1136"""
    return np.abs(np.sum(np.conj(g) * g, axis=1) * ipm)


def norm_within_radius(g, r, radius, ipm):
    """mesh/data.py:419-422"""
    m = np.where(r[None, :] <= radius, g, 0) if g.ndim == 2 else np.where(r <= radius, g, 0)
    return norm(m, ipm)


def r_expectation(g, r, ipm):
    """expectation of ElementWiseMultiplyOperator(r_mesh) (mesh_operators.py:1082-1084)."""
    rr = r[None, :] if g.ndim == 2 else r
    return float(np.real(np.sum(np.conj(g) * (rr * g)) * ipm))


def sh_z_expectation(g, c_l, r, ipm):
    """<z>: tridiagonal-in-l operator r_j c_l (mesh_operators.py:1086-1104)."""
    w = c_l[:, None] * r[None, :]
    zg = np.zeros_like(g)
    zg[:-1] += w * g[1:]
    zg[1:] += w * g[:-1]
    return float(np.real(np.sum(np.conj(g) * zg) * ipm))


def h0_expectation(g, h_diag, h_off, ipm):
    """<H0> (mesh/meshes.py:219-223)."""
    g2 = np.atleast_2d(g)
    hg = np.atleast_2d(h_diag) * g2
    hg[:, 1:] += h_off * g2[:, :-1]
    hg[:, :-1] += h_off * g2[:, 1:]
    return float(np.real(np.sum(np.conj(g2) * hg) * ipm))


def sh_len_total_energy_expectation(g, h_diag, h_off, c_l, x_j, efield, ipm):
    """<H0 + Hint>, length gauge (mesh/meshes.py:225-229, mesh_operators.py:1020-1035);
    efield = E(t + dt/2) as the reference samples it (:1011-1013)."""
    w = efield * c_l[:, None] * x_j[None, :]
    ig = np.zeros_like(g)
    ig[:-1] += w * g[1:]
    ig[1:] += w * g[:-1]
    return h0_expectation(g, h_diag, h_off, ipm) + float(np.real(np.sum(np.conj(g) * ig) * ipm))


# --------------------------------------------------------------------------
# whole-run drivers (a13: the loop of mesh/sims.py:289-321)
# --------------------------------------------------------------------------
def run_sh(problem, *, store_every_step=True, state_l=None, state_rows=None):
    """Run a dumped SphericalHarmonic problem (see oracle/make_golden.py for the keys).
    Returns dict(g=final g, norm=[...], inner_products=[n_data, n_states])."""
    g = np.array(problem["g0"], dtype=np.complex128)
    kind = str(problem["kind"])
    taus = problem["taus"]
    fields = problem["fields"]
    ipm = float(problem["delta_r"])
    mask = problem["mask"]
    norms = [norm(g, ipm)]
    ips = []
    if state_rows is not None:
        ips.append(inner_product_rows(g, state_l, state_rows, ipm))
    for n in range(len(taus)):
        if kind == "sh_len_so":
            g = sh_len_so_step(g, problem["h_diag"], problem["h_off"], problem["c_l"], problem["x_j"], mask, taus[n], fields[n])
        elif kind == "sh_vel_so":
            g = sh_vel_so_step(
                g, problem["h_diag"], problem["h_off"], problem["c_l"], problem["f1_l"], problem["y_j"], problem["z_j"], mask, taus[n], fields[n]
            )
        elif kind == "sh_len_adi":
            g = sh_len_adi_step(g, problem["h_diag"], problem["h_off"], problem["c_l"], problem["x_j"], mask, taus[n], fields[n])
        else:
            raise ValueError(kind)
        if store_every_step or n == len(taus) - 1:
            norms.append(norm(g, ipm))
            if state_rows is not None:
                ips.append(inner_product_rows(g, state_l, state_rows, ipm))
    return dict(g=g, norm=np.array(norms), inner_products=np.array(ips))


def run_line(problem, *, store_every_step=True, state_rows=None):
    g = np.array(problem["g0"], dtype=np.complex128)
    kind = str(problem["kind"])
    taus = problem["taus"]
    fields = problem["fields"]
    ipm = float(problem["delta_z"])
    mask = problem["mask"]
    norms = [norm(g, ipm)]
    ips = []
    if state_rows is not None:
        ips.append(np.array([inner_product_full(s, g, ipm) for s in state_rows]))
    for n in range(len(taus)):
        if kind == "line_len_cn":
            g = line_cn_len_step(g, problem["h_diag"], problem["h_off"], problem["w_z"], mask, taus[n], fields[n])
        elif kind == "line_len_so":
            g = line_so_len_step(g, problem["h_diag"], problem["h_off"], problem["w_z"], mask, taus[n], fields[n])
        elif kind == "line_vel_so":
            g = line_so_vel_step(g, problem["h_diag"], problem["h_off"], float(problem["v_pref"]), mask, taus[n], fields[n])
        else:
            raise ValueError(kind)
        if store_every_step or n == len(taus) - 1:
            norms.append(norm(g, ipm))
            if state_rows is not None:
                ips.append(np.array([inner_product_full(s, g, ipm) for s in state_rows]))
    return dict(g=g, norm=np.array(norms), inner_products=np.array(ips))
