"""ctypes wrapper over oracle/csrc/jk_oracle.c (the C restatement).  TEST INFRASTRUCTURE."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libjk_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    if force or not os.path.exists(_LIB):
        if force and os.path.exists(_LIB):
            os.remove(_LIB)
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.sqcc_oracle_num_threads.restype = ctypes.c_int
        _lib.sqcc_oracle_synth_canonical.argtypes = [ctypes.c_int64, ctypes.c_uint64, _dp]
        _lib.sqcc_oracle_jk_entries.argtypes = [ctypes.c_int, _dp, ctypes.c_int64, ctypes.c_uint64, _dp,
                                                ctypes.c_int, _ip, ctypes.c_int64, _dp, _dp]
        _lib.sqcc_oracle_jk_packed.argtypes = [_dp, ctypes.c_int64, _dp, ctypes.c_int, ctypes.c_int, _dp, _dp]
        _lib.sqcc_oracle_synth_rows.argtypes = [ctypes.c_int64, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, _dp]
        _lib.sqcc_oracle_jk_packed_rows.argtypes = [_dp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _dp,
                                                    ctypes.c_int, ctypes.c_int, _dp, _dp]
        _lib.sqcc_oracle_mp2_entries_synth.argtypes = [ctypes.c_int64, ctypes.c_uint64, _dp, ctypes.c_int64,
                                                       ctypes.c_int64, _ip, _ip, ctypes.c_int64, _dp]
        _lib.sqcc_oracle_synth_slabs_rs.argtypes = [ctypes.c_int64, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, _dp]
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def num_threads():
    return lib().sqcc_oracle_num_threads()


def synth_canonical(n, seed):
    p = n * (n + 1) // 2
    out = np.empty(p * (p + 1) // 2)
    lib().sqcc_oracle_synth_canonical(n, seed, _p(out))
    return out


SRC_FULL, SRC_PACKED, SRC_SYNTH = 0, 1, 2


def jk_entries(src, buf, n, seed, d, entries=None, want_k=True):
    """J/K by the reference loop nest (hf_ksdft.py:483-495) at `entries` [(p,q),...] or all N*N."""
    d3 = np.ascontiguousarray(d if d.ndim == 3 else d[None], dtype=np.float64)
    nspin = d3.shape[0]
    if entries is None:
        cnt, pq = n * n, None
    else:
        pq = np.ascontiguousarray(np.asarray(entries, dtype=np.int64).reshape(-1, 2))
        cnt = len(pq)
    j = np.zeros(cnt)
    k = np.zeros((nspin, cnt)) if want_k else None
    b = np.ascontiguousarray(buf, dtype=np.float64) if buf is not None else None
    lib().sqcc_oracle_jk_entries(src, _p(b), n, seed, _p(d3), nspin,
                                 pq.ctypes.data_as(_ip) if pq is not None else None, cnt, _p(j), _p(k))
    if entries is None:
        j = j.reshape(n, n)
        if want_k:
            k = k.reshape(nspin, n, n)
            k = k[0] if d.ndim == 2 else k
    elif want_k and d.ndim == 2:
        k = k[0]
    return j, k


def jk_packed(packed, n, d, want_k=True):
    """Whole J/K from the canonical packed vector by the 8-fold formula (multi-core port)."""
    d3 = np.ascontiguousarray(d if d.ndim == 3 else d[None], dtype=np.float64)
    nspin = d3.shape[0]
    j = np.zeros((n, n))
    k = np.zeros((nspin, n, n)) if want_k else None
    lib().sqcc_oracle_jk_packed(_p(np.ascontiguousarray(packed)), n, _p(d3), nspin, int(want_k), _p(j), _p(k))
    if want_k and d.ndim == 2:
        k = k[0]
    return j, k


def synth_rows(n, seed, ij0, ij1):
    """Canonical packed rows [ij0, ij1) of the synthetic tensor, back to back."""
    out = np.empty(ij1 * (ij1 + 1) // 2 - ij0 * (ij0 + 1) // 2)
    lib().sqcc_oracle_synth_rows(n, seed, ij0, ij1, _p(out))
    return out


def jk_packed_rows(rows, n, ij0, ij1, d, want_k=True):
    """Partial (symmetrised) J/K contributed by packed rows [ij0, ij1) -- what one pair-block shard computes."""
    d3 = np.ascontiguousarray(d if d.ndim == 3 else d[None], dtype=np.float64)
    nspin = d3.shape[0]
    j = np.zeros((n, n))
    k = np.zeros((nspin, n, n)) if want_k else None
    lib().sqcc_oracle_jk_packed_rows(_p(np.ascontiguousarray(rows)), n, ij0, ij1, _p(d3), nspin, int(want_k), _p(j), _p(k))
    if want_k and d.ndim == 2:
        k = k[0]
    return j, k


def jk_packed_segments(n, seed, segments, d, want_k=True):
    """Partial J/K of a shard given as a list of canonical row ranges [(ij0, ij1), ...] of the synthetic tensor
    (a tile-range shard of sqcc_shard_tiles is a union of such runs)."""
    nspin = 1 if d.ndim == 2 else 2
    j = np.zeros((n, n))
    k = np.zeros((nspin, n, n)) if want_k else None
    for ij0, ij1 in segments:
        rows = synth_rows(n, seed, ij0, ij1)
        jp, kp = jk_packed_rows(rows, n, ij0, ij1, d, want_k)
        j += jp
        if want_k:
            k += kp if d.ndim == 3 else kp[None]
    if want_k and d.ndim == 2:
        k = k[0]
    return j, k


def mp2_entries_synth(n, seed, c, i, j, ab):
    """(ia|jb) of the synthetic tensor for one occupied MO pair (i, j) and MO column pairs ab = [(a, b), ...], by the
    reference's four nested np.dot steps (mp2.py:123-170).  `c` is the full [N,N] AO x MO matrix; a, b are MO columns."""
    c = np.ascontiguousarray(c, dtype=np.float64)
    ab = np.asarray(ab, dtype=np.int64).reshape(-1, 2)
    ca, cb = np.ascontiguousarray(ab[:, 0]), np.ascontiguousarray(ab[:, 1])
    out = np.zeros(len(ab))
    lib().sqcc_oracle_mp2_entries_synth(n, seed, _p(c), int(i), int(j), ca.ctypes.data_as(_ip), cb.ctypes.data_as(_ip),
                                        len(ab), _p(out))
    return out


def synth_slabs_rs(n, seed, rs0, rs1):
    """ERI[:, :, r, s] for the canonical pairs rs in [rs0, rs1) of the synthetic tensor: [rs1 - rs0, N, N]."""
    out = np.empty((rs1 - rs0, n, n))
    lib().sqcc_oracle_synth_slabs_rs(n, seed, rs0, rs1, _p(out))
    return out
