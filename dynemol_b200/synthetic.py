"""Synthetic EHT-structured workloads for the propagator.

The reference builds S (STO overlaps, overlap_D.f) and h (hamiltonians.f) on the
host; that assembly is out of scope and stays on the reference path.  For
benchmarks and parity tests we need inputs with the same *structure*:

  * n_atoms = N/4 atoms on a jittered cubic lattice (a = 2.9 A, sigma = 0.15 A,
    seed 1234+N), four orbitals per atom (s, px, py, pz), species alternating
    "S-like" (IP -20.0 / -11.0 eV) and "Li-like" (IP -5.4 / -3.5 eV),
    k_WH = 1.75, V_shift = 0;
  * S = exact overlap (Gram) matrix of normalised s/p Gaussians of exponent
    zeta, hence symmetric positive definite with unit diagonal, set to zero
    beyond cutoff_Angs = 12 A (constants_m.f:34);
  * h_ij = X_ij * S_ij with X_ij exactly as hamiltonians.f:46-61;
  * packets: electron = normalised random real vector on the first 64 orbitals
    (seed 42), hole on the next 64 (seed 43); Psi_ket = C, Psi_bra = S C
    (ElHl_Chebyshev.f:126-129); fragments = 4 equal contiguous orbital blocks.

Overlap formulas (R = A - B, E = exp(-zeta R^2 / 2), all orbitals share zeta):
  <sA|sB> = E ; <sA|pB_b> = sqrt(zeta) R_b E ; <pA_a|sB> = -sqrt(zeta) R_a E ;
  <pA_a|pB_b> = (delta_ab - zeta R_a R_b) E.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

H_BAR = 6.58264e-4        # eV*ps (constants_m.f:23)
CUTOFF_ANGS = 12.0        # constants_m.f:34
LATTICE_A = 2.9
JITTER = 0.15
ZETA = 0.30               # A^-2; nearest-neighbour s-s overlap ~ 0.28, cond(S) ~ 1e2
K_WH = 1.75
IP_S_LIKE = (-20.0, -11.0)
IP_LI_LIKE = (-5.4, -3.5)


@dataclass
class Workload:
    N: int
    S: np.ndarray            # (N,N) float64, Fortran order
    h: np.ndarray            # (N,N) float64, Fortran order
    IP: np.ndarray
    k_WH: np.ndarray
    V_shift: np.ndarray
    C: np.ndarray            # (N,2) real packet coefficients (el, hl)
    Psi_bra: np.ndarray      # (N,2) complex128, Fortran order  = S C
    Psi_ket: np.ndarray      # (N,2) complex128, Fortran order  = C
    fragment: np.ndarray     # (N,) int32 in 0..3
    positions: np.ndarray


def lattice(n_atoms: int, seed: int):
    """Jittered simple-cubic positions and the S-like/Li-like parity of each site."""
    L = int(np.ceil(n_atoms ** (1.0 / 3.0) - 1e-9))
    idx = np.array([(i, j, k) for i in range(L) for j in range(L) for k in range(L)][:n_atoms], dtype=np.float64)
    rng = np.random.default_rng(seed)
    pos = idx * LATTICE_A + rng.normal(0.0, JITTER, size=idx.shape)
    species = (idx.sum(axis=1).astype(np.int64) % 2).astype(np.int64)   # 0 = S-like, 1 = Li-like
    return pos, species


def li2s_lattice(nx: int, ny: int, nz: int, a: float = 5.72, jitter: float = 0.05, seed: int = 2592):
    """Anti-fluorite Li2S supercell (BASELINE config 1, examples/Li2S-crystal: 1728 Li + 864 S = 2592 atoms for
    6x6x6 cells, N = 10368 orbitals with this 4-orbital model).  S on the fcc sites, Li on the eight (1/4,1/4,1/4)-type
    sites of every conventional cell; small thermal jitter.  species: 0 = S-like, 1 = Li-like."""
    fcc = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
    tet = np.array([[x, y, z] for x in (.25, .75) for y in (.25, .75) for z in (.25, .75)])
    cell = np.concatenate([fcc, tet]); sp = np.array([0] * 4 + [1] * 8)
    shifts = np.array([[i, j, k] for i in range(nx) for j in range(ny) for k in range(nz)], dtype=np.float64)
    pos = (shifts[:, None, :] + cell[None, :, :]).reshape(-1, 3) * a
    species = np.tile(sp, shifts.shape[0])
    rng = np.random.default_rng(seed)
    return pos + rng.normal(0.0, jitter, size=pos.shape), species


def orbital_params(species: np.ndarray):
    n_atoms = species.shape[0]
    IP = np.empty(4 * n_atoms)
    ip_s = np.where(species == 0, IP_S_LIKE[0], IP_LI_LIKE[0])
    ip_p = np.where(species == 0, IP_S_LIKE[1], IP_LI_LIKE[1])
    IP[0::4] = ip_s
    IP[1::4] = ip_p; IP[2::4] = ip_p; IP[3::4] = ip_p
    k_WH = np.full(4 * n_atoms, K_WH)
    V_shift = np.zeros(4 * n_atoms)
    return IP, k_WH, V_shift


def overlap_numpy(pos: np.ndarray, zeta: float = ZETA) -> np.ndarray:
    n_atoms = pos.shape[0]
    R = pos[:, None, :] - pos[None, :, :]                # R[A,B,:] = A - B
    R2 = np.einsum("abk,abk->ab", R, R)
    E = np.exp(-0.5 * zeta * R2)
    E[R2 > CUTOFF_ANGS ** 2] = 0.0
    sz = np.sqrt(zeta)
    S = np.zeros((n_atoms, 4, n_atoms, 4))
    S[:, 0, :, 0] = E
    for b in range(3):
        S[:, 0, :, 1 + b] = sz * R[:, :, b] * E          # <sA|pB_b>
        S[:, 1 + b, :, 0] = -sz * R[:, :, b] * E         # <pA_b|sB>
        for a in range(3):
            S[:, 1 + a, :, 1 + b] = ((1.0 if a == b else 0.0) - zeta * R[:, :, a] * R[:, :, b]) * E
    S = S.reshape(4 * n_atoms, 4 * n_atoms)
    S = 0.5 * (S + S.T)
    np.fill_diagonal(S, 1.0)
    return np.asfortranarray(S)


def x_matrix(IP, k_WH, V_shift):
    """X_ij of hamiltonians.f:33-63 for all pairs (vectorised)."""
    c1 = IP[:, None] - IP[None, :]
    c2 = IP[:, None] + IP[None, :]
    c3 = (c1 / c2) * (c1 / c2)
    c4 = (V_shift[:, None] + V_shift[None, :]) * 0.5
    kwh = (k_WH[:, None] + k_WH[None, :]) * 0.5
    k_eff = kwh + c3 + c3 * c3 * (1.0 - kwh)
    X = k_eff * c2 * 0.5 + c4
    X[np.diag_indices_from(X)] = IP + V_shift
    return X


def packets(S: np.ndarray, N: int):
    """Electron on the first block of orbitals, hole on the next one; C^T S C = 1."""
    w = min(64, max(1, N // 2))
    C = np.zeros((N, 2))
    C[0:w, 0] = np.random.default_rng(42).normal(size=w)
    C[w:2 * w, 1] = np.random.default_rng(43).normal(size=w)
    for p in range(2):
        nrm = float(C[:, p] @ (S @ C[:, p]))
        C[:, p] /= np.sqrt(nrm)
    Psi_ket = np.asfortranarray(C.astype(np.complex128))
    Psi_bra = np.asfortranarray((S @ C).astype(np.complex128))
    return C, Psi_bra, Psi_ket


def fragments(N: int) -> np.ndarray:
    return np.minimum(3, (np.arange(N) * 4) // N).astype(np.int32)


def make_workload(N: int, zeta: float = ZETA, seed: int | None = None) -> Workload:
    """CPU/numpy generator (fine up to N ~ 8k)."""
    assert N % 4 == 0, "4 orbitals per atom"
    pos, species = lattice(N // 4, 1234 + N if seed is None else seed)
    IP, k_WH, V_shift = orbital_params(species)
    S = overlap_numpy(pos, zeta)
    h = np.asfortranarray(x_matrix(IP, k_WH, V_shift) * S)
    C, Psi_bra, Psi_ket = packets(S, N)
    return Workload(N, S, h, IP, k_WH, V_shift, C, Psi_bra, Psi_ket, fragments(N), pos)


def perturb_positions(pos: np.ndarray, step: int, amplitude: float = 0.01) -> np.ndarray:
    """Deterministic small nuclear motion for multi-step trajectories."""
    rng = np.random.default_rng(9000 + step)
    return pos + rng.normal(0.0, amplitude, size=pos.shape)


def workload_at(pos: np.ndarray, species: np.ndarray, zeta: float = ZETA):
    """S and h for a given geometry (used for trajectories: geometry changes per step)."""
    IP, k_WH, V_shift = orbital_params(species)
    S = overlap_numpy(pos, zeta)
    h = np.asfortranarray(x_matrix(IP, k_WH, V_shift) * S)
    return S, h


# ----------------------------------------------------------------------------- device-side generation (bench only)
def overlap_rows_torch(pos, a0: int, a1: int, zeta: float = ZETA):
    """Rows 4*a0 .. 4*a1-1 of S (all N columns) on pos.device, float64, shape (4*(a1-a0), N)."""
    import torch
    n_atoms = pos.shape[0]
    sz = float(np.sqrt(zeta))
    eye3 = torch.eye(3, device=pos.device, dtype=torch.float64)
    R = pos[a0:a1, None, :] - pos[None, :, :]
    R2 = (R * R).sum(-1)
    E = torch.exp(-0.5 * zeta * R2)
    E = torch.where(R2 > CUTOFF_ANGS ** 2, torch.zeros_like(E), E)
    blk = torch.empty((a1 - a0, 4, n_atoms, 4), device=pos.device, dtype=torch.float64)
    blk[:, 0, :, 0] = E
    blk[:, 0, :, 1:] = sz * R * E[..., None]
    blk[:, 1:, :, 0] = (-sz * R * E[..., None]).permute(0, 2, 1)
    pp = (eye3[None, None] - zeta * R[..., :, None] * R[..., None, :]) * E[..., None, None]   # [A,B,a,b]
    blk[:, 1:, :, 1:] = pp.permute(0, 2, 1, 3)
    out = blk.reshape(4 * (a1 - a0), 4 * n_atoms)
    # exact unit diagonal / symmetric in exact arithmetic; enforce the diagonal like overlap_numpy does
    idx = torch.arange(4 * a0, 4 * a1, device=pos.device)
    out[idx - 4 * a0, idx] = 1.0
    return out


def huckel_rows_torch(S_rows, r0: int, IPt, Kt, Vt):
    """h rows r0.. for the given S rows: h_ij = X_ij S_ij (hamiltonians.f:33-63)."""
    r1 = r0 + S_rows.shape[0]
    c1 = IPt[r0:r1, None] - IPt[None, :]
    c2 = IPt[r0:r1, None] + IPt[None, :]
    c3 = (c1 / c2) * (c1 / c2)
    c4 = (Vt[r0:r1, None] + Vt[None, :]) * 0.5
    kwh = (Kt[r0:r1, None] + Kt[None, :]) * 0.5
    k_eff = kwh + c3 + c3 * c3 * (1.0 - kwh)
    h = (k_eff * c2 * 0.5 + c4) * S_rows
    import torch
    idx = torch.arange(r0, r1, device=S_rows.device)
    h[idx - r0, idx] = IPt[r0:r1] + Vt[r0:r1]        # X_ii * S_ii, S_ii = 1
    return h


def S_h_torch_from_positions(pos_np, species, device, zeta: float = ZETA, row_block: int = 512):
    """S and h on the GPU for an arbitrary geometry (same recipe as overlap_numpy / x_matrix)."""
    import torch
    IP, k_WH, V_shift = orbital_params(species)
    pos = torch.tensor(pos_np, device=device, dtype=torch.float64)
    n_atoms = pos.shape[0]; N = 4 * n_atoms
    IPt = torch.tensor(IP, device=device); Kt = torch.tensor(k_WH, device=device); Vt = torch.tensor(V_shift, device=device)
    S = torch.empty((N, N), device=device, dtype=torch.float64)
    h = torch.empty_like(S)
    for a0 in range(0, n_atoms, row_block):
        a1 = min(n_atoms, a0 + row_block)
        S[4 * a0:4 * a1, :] = overlap_rows_torch(pos, a0, a1, zeta)
    S = 0.5 * (S + S.T)
    S.fill_diagonal_(1.0)
    for a0 in range(0, n_atoms, row_block):
        a1 = min(n_atoms, a0 + row_block)
        h[4 * a0:4 * a1, :] = huckel_rows_torch(S[4 * a0:4 * a1, :], 4 * a0, IPt, Kt, Vt)
    return S, h, dict(pos=pos_np, species=species, IP=IP, k_WH=k_WH, V_shift=V_shift)


def make_S_h_torch(N: int, device, zeta: float = ZETA, row_block: int = 512):
    """The jittered-cubic-lattice workload of make_workload() built on the GPU with torch (input generation is
    plumbing, not the product): (symmetric) S and h as torch float64 (N,N) tensors plus the numpy metadata."""
    pos_np, species = lattice(N // 4, 1234 + N)
    return S_h_torch_from_positions(pos_np, species, device, zeta, row_block)


def make_h_shard_colmajor_torch(N: int, row0: int, n_rows: int, device, zeta: float = ZETA, row_block: int = 256,
                                dense_tail: bool = True):
    """Rows row0..row0+n_rows-1 of the H' SURROGATE of the N=65536 throughput configs (all N columns), laid out
    COLUMN-major with leading dimension n_rows, i.e. as a torch tensor of shape (N, n_rows).

    Surrogate = Hueckel matrix h (exact recipe above) + a dense, exponentially decaying tail
    1e-3 * exp(-R_AB / 10 A) * u_ij, u_ij ~ U(-1,1) seeded by the row block.  The true H' = S^-1 h is dense with
    full-precision mantissas everywhere; h alone is 98% exact zeros at this size, which lowers the FP64 datapath
    power enough to lift the board off its power cap and would flatter the measured bandwidth.  Only throughput is
    measured on this operator; forming S^-1 h at N=65536 needs a distributed factorisation."""
    import torch
    assert row0 % 4 == 0 and n_rows % 4 == 0
    pos_np, species = lattice(N // 4, 1234 + N)
    IP, k_WH, V_shift = orbital_params(species)
    pos = torch.tensor(pos_np, device=device, dtype=torch.float64)
    IPt = torch.tensor(IP, device=device); Kt = torch.tensor(k_WH, device=device); Vt = torch.tensor(V_shift, device=device)
    out = torch.empty((N, n_rows), device=device, dtype=torch.float64)
    for a0 in range(row0 // 4, (row0 + n_rows) // 4, row_block):
        a1 = min((row0 + n_rows) // 4, a0 + row_block)
        S_rows = overlap_rows_torch(pos, a0, a1, zeta)
        h_rows = huckel_rows_torch(S_rows, 4 * a0, IPt, Kt, Vt)
        if dense_tail:
            g = torch.Generator(device=device); g.manual_seed(777 + a0)
            R = (pos[a0:a1, None, :] - pos[None, :, :]).norm(dim=-1)                    # (atoms_blk, n_atoms)
            env = (1.0e-3 * torch.exp(-R / 10.0)).repeat_interleave(4, 0).repeat_interleave(4, 1)
            u = torch.rand(h_rows.shape, generator=g, device=device, dtype=torch.float64) * 2.0 - 1.0
            h_rows = h_rows + env * u
        out[:, 4 * a0 - row0:4 * a1 - row0] = h_rows.t()
    return out
