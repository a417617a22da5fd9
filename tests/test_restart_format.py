"""CPU suite: Security_copy.dat / Restart_copy.dat (backup.f:329-445) -- Fortran unformatted record framing and the
exact record sequence of Security_Copy_Cheb, written and read back."""
import struct

import numpy as np

from dynemol_b200.restart import ChebState, read_restart_copy, write_security_copy


def _state(N=12, n_atoms=3, seed=0):
    rng = np.random.default_rng(seed)
    z = lambda: np.asfortranarray(rng.normal(size=(N, 2)) + 1j * rng.normal(size=(N, 2)))
    return ChebState(frame=7, it=8, t=3.5e-3, eh_tag=["el", "hl"], DUAL_bra=z(), DUAL_ket=z(), AO_bra=z(), AO_ket=z(),
                     Net_Charge=rng.normal(size=n_atoms))


def test_record_sequence_matches_security_copy_cheb(tmp_path):
    st = _state()
    path = tmp_path / "Security_copy.dat"
    write_security_copy(str(path), st)
    raw = path.read_bytes()
    # walk the records: [len][payload][len] ...
    recs, off = [], 0
    while off < len(raw):
        (n,) = struct.unpack_from("<i", raw, off)
        payload = raw[off + 4: off + 4 + n]
        assert struct.unpack_from("<i", raw, off + 4 + n)[0] == n
        recs.append(payload); off += 8 + n
    N, n_part, n_atoms = 12, 2, 3
    assert [len(r) for r in recs[:7]] == [4, 4, 8, 4, 4, 4, 2 * n_part]          # frame, it, t, sizes, eh_tag
    assert struct.unpack("<i", recs[0])[0] == 7 and struct.unpack("<i", recs[1])[0] == 8
    assert struct.unpack("<d", recs[2])[0] == 3.5e-3 and recs[6] == b"elhl"
    assert len(recs) == 7 + 2 * n_part + 1
    for r in recs[7:7 + 2 * n_part]:
        assert len(r) == N * 2 * 16                                                # (bra(i), ket(i)) pairs of complex*16
    first = np.frombuffer(recs[7], dtype="<c16").reshape(N, 2)
    assert np.array_equal(first[:, 0], st.DUAL_bra[:, 0]) and np.array_equal(first[:, 1], st.DUAL_ket[:, 0])
    assert len(recs[-1]) == 8 * n_atoms * n_atoms                                  # the implied-DO quirk of backup.f:390


def test_round_trip(tmp_path):
    st = _state(N=40, n_atoms=10, seed=3)
    path = tmp_path / "Restart_copy.dat"
    write_security_copy(str(path), st)
    back = read_restart_copy(str(path))
    assert (back.frame, back.it, back.t, back.eh_tag) == (st.frame, st.it, st.t, st.eh_tag)
    for name in ("DUAL_bra", "DUAL_ket", "AO_bra", "AO_ket", "Net_Charge"):
        assert np.array_equal(getattr(back, name), getattr(st, name)), name


def test_long_records_are_split_into_subrecords(tmp_path, monkeypatch):
    """Records beyond 2^31 - 9 bytes (Net_Charge at >= 16384 atoms: n_atoms^2 * 8 B) are written as subrecords the way
    the gfortran / ifort runtimes do it: negative leading marker = "continued", negative trailing marker = "has a
    predecessor".  Exercised with a tiny limit."""
    import dynemol_b200.restart as R
    monkeypatch.setattr(R, "MAX_SUBRECORD", 100)
    st = _state(N=6, n_atoms=7, seed=5)                       # Net_Charge record: 7*7*8 = 392 B -> 100+100+100+92
    path = tmp_path / "Security_copy.dat"
    R.write_security_copy(str(path), st)
    raw = path.read_bytes()
    tail = raw[-(392 + 4 * 8):]                               # the last record: 4 subrecords, 8 marker bytes each
    subs, off = [], 0
    while off < len(tail):
        (lead,) = struct.unpack_from("<i", tail, off)
        n = abs(lead)
        (trail,) = struct.unpack_from("<i", tail, off + 4 + n)
        subs.append((lead, trail)); off += 8 + n
    assert subs == [(-100, 100), (-100, -100), (-100, -100), (92, -92)]
    back = R.read_restart_copy(str(path))
    assert np.array_equal(back.Net_Charge, st.Net_Charge) and np.array_equal(back.AO_ket, st.AO_ket)
    # the interleaved packet records (6*2*16 = 192 B) were split as well and read back intact
    assert np.array_equal(back.DUAL_bra, st.DUAL_bra)
