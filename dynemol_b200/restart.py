"""Security_copy.dat / Restart_copy.dat of the slice_Cheb driver (reference: backup.f:329-445,
Security_Copy_Cheb / Restart_State_Cheb), byte-compatible with what the Intel/GNU Fortran runtime writes for
`form="unformatted"` sequential files: every WRITE is one record framed by a 4-byte little-endian length before and
after the payload.  Records, in order (backup.f:372-392):

    frame (int32) | it (int32) | t (real*8) | basis_size (int32) | n_part (int32) | size(eh_tag) (int32)
    eh_tag(1:n_part)                       character(len=2) each
    per particle j:  (DUAL_bra(i,j), DUAL_ket(i,j), i=1,basis_size)     complex*16 pairs, interleaved
                     (AO_bra(i,j),   AO_ket(i,j),   i=1,basis_size)
    (Net_Charge, i=1,size(Net_Charge))     the WHOLE array repeated size(Net_Charge) times (the implied-DO of the
                                           reference writes the array once per index; kept for compatibility)

The propagator itself needs no device checkpoint: its state between calls is host-side except
save_tau, which the reference does not store either (after a restart the first step starts again from tau_max,
ElHl_Chebyshev.f:34-38,182).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np


@dataclass
class ChebState:
    frame: int
    it: int
    t: float
    eh_tag: list
    DUAL_bra: np.ndarray      # (N, n_part) complex128
    DUAL_ket: np.ndarray
    AO_bra: np.ndarray
    AO_ket: np.ndarray
    Net_Charge: np.ndarray    # (n_atoms,) float64


# Records longer than 2^31 - 9 bytes are split into subrecords by the gfortran / ifort runtimes: every subrecord is framed
# by its own 4-byte markers holding its length; the LEADING marker is negative when another subrecord follows, the
# TRAILING marker is negative when a subrecord precedes.  (Net_Charge is written size(Net_Charge) times, backup.f:392:
# n_atoms^2 * 8 bytes, i.e. exactly 2 GiB at 16384 atoms.)
MAX_SUBRECORD = 2147483639


def _rec(f, payload: bytes):
    _rec_stream(f, len(payload), iter((payload,)))


def _rec_stream(f, total: int, pieces):
    """One Fortran record of `total` bytes whose payload arrives as an iterable of bytes objects (never materialised whole)."""
    buf = b""
    remaining = total
    first = True
    while True:
        sub = min(remaining, MAX_SUBRECORD)
        more = remaining > sub
        f.write(struct.pack("<i", -sub if more else sub))
        left = sub
        while left > 0:
            if not buf:
                buf = next(pieces)
            take = buf[:left]
            f.write(take)
            left -= len(take)
            buf = buf[len(take):]
        f.write(struct.pack("<i", sub if first else -sub))
        remaining -= sub
        first = False
        if not more:
            break


def write_security_copy(path: str, st: ChebState) -> None:
    """backup.f:329-394 Security_Copy_Cheb."""
    N, n_part = st.AO_bra.shape
    with open(path, "wb") as f:
        _rec(f, struct.pack("<i", st.frame))
        _rec(f, struct.pack("<i", st.it))
        _rec(f, struct.pack("<d", st.t))
        _rec(f, struct.pack("<i", N))
        _rec(f, struct.pack("<i", n_part))
        _rec(f, struct.pack("<i", len(st.eh_tag)))
        _rec(f, b"".join(t.encode("ascii").ljust(2)[:2] for t in st.eh_tag[:n_part]))
        for j in range(n_part):
            for a, b in ((st.DUAL_bra, st.DUAL_ket), (st.AO_bra, st.AO_ket)):
                inter = np.empty((N, 2), dtype=np.complex128)
                inter[:, 0] = a[:, j]; inter[:, 1] = b[:, j]
                _rec(f, inter.astype("<c16").tobytes())
        nc = np.ascontiguousarray(st.Net_Charge, dtype="<f8")
        one = nc.tobytes()
        _rec_stream(f, len(one) * nc.size, (one for _ in range(nc.size)))      # the array, size(Net_Charge) times, streamed


def _read_rec(f, keep: int | None = None):
    """One Fortran record (all its subrecords).  keep=None: returns the payload.  keep=k: returns (first k bytes, total
    length) without holding the rest in memory (the Net_Charge record can be 2 GiB of repeats)."""
    out = []
    kept = 0
    total = 0
    first = True
    while True:
        head = f.read(4)
        if len(head) != 4:
            raise EOFError("truncated Fortran record")
        (n,) = struct.unpack("<i", head)
        sub = abs(n)
        if keep is None:
            data = f.read(sub)
            if len(data) != sub:
                raise ValueError("corrupt Fortran record framing")
            out.append(data)
        else:
            take = min(sub, keep - kept)
            if take > 0:
                out.append(f.read(take)); kept += take
            f.seek(sub - take, 1)
        total += sub
        tail = f.read(4)
        if len(tail) != 4:
            raise ValueError("corrupt Fortran record framing")
        (m,) = struct.unpack("<i", tail)
        if abs(m) != sub or (m < 0) == first:
            raise ValueError("corrupt Fortran record framing")
        first = False
        if n >= 0:
            break
    payload = b"".join(out)
    return payload if keep is None else (payload, total)


def read_restart_copy(path: str) -> ChebState:
    """backup.f:399-445 Restart_State_Cheb."""
    with open(path, "rb") as f:
        frame = struct.unpack("<i", _read_rec(f))[0]
        it = struct.unpack("<i", _read_rec(f))[0]
        t = struct.unpack("<d", _read_rec(f))[0]
        N = struct.unpack("<i", _read_rec(f))[0]
        n_part = struct.unpack("<i", _read_rec(f))[0]
        n_tag = struct.unpack("<i", _read_rec(f))[0]
        raw = _read_rec(f)
        tags = [raw[2 * i:2 * i + 2].decode("ascii") for i in range(len(raw) // 2)]
        arrs = [np.empty((N, n_part), dtype=np.complex128, order="F") for _ in range(4)]
        for j in range(n_part):
            for k in (0, 2):
                inter = np.frombuffer(_read_rec(f), dtype="<c16").reshape(N, 2)
                arrs[k][:, j] = inter[:, 0]; arrs[k + 1][:, j] = inter[:, 1]
        pos = f.tell()
        _, total = _read_rec(f, keep=0)
        n_atoms = int(round(np.sqrt(total // 8)))
        if n_atoms * n_atoms * 8 != total:
            raise ValueError("Net_Charge record is not size(Net_Charge) copies of the array")
        f.seek(pos)
        head, _ = _read_rec(f, keep=8 * n_atoms)
        nc = np.frombuffer(head, dtype="<f8").copy()
    assert n_tag >= n_part
    return ChebState(frame, it, t, tags, arrs[0], arrs[1], arrs[2], arrs[3], nc)
