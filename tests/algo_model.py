"""numpy model of the multi-pass packed-real FFT correlation the CUDA kernels
implement (TEST INFRASTRUCTURE -- documents and checks the index algebra of
DESIGN.md section 3; it is neither the oracle nor a product path).

Layout: a length-L real signal (L = 2M) is packed as M complex points
z[m] = x[2m] + i x[2m+1].  M = N_1 * ... * N_p.  Pass a transforms axis a of
the C-ordered array (N_1, ..., N_p) and multiplies by W_{Q_a}^{k_a * m'_a}
(Q_a = N_a...N_p, m'_a the flattened index of the trailing axes).  The
spectrum is left in "digit-transposed" order: address ((k_1 N_2 + k_2) N_3 +
...) holds frequency k = k_1 + N_1 k_2 + N_1 N_2 k_3 + ...  The inverse runs
the passes backwards, so no transposition is ever materialised.
"""
import numpy as np


def forward_packed(x, dims):
    """x real (n,), zero-padded implicitly to L = 2*prod(dims).  Returns Z in
    digit-transposed order, flat (M,)."""
    M = int(np.prod(dims))
    n = len(x)
    assert M >= n
    xp = np.zeros(2 * M)
    xp[:n] = x
    z = (xp[0::2] + 1j * xp[1::2]).reshape(dims)
    p = len(dims)
    for a in range(p):
        z = np.fft.fft(z, axis=a)
        if a < p - 1:
            Q = int(np.prod(dims[a:]))
            mrest = np.arange(Q // dims[a]).reshape(dims[a + 1:])
            ka = np.arange(dims[a]).reshape((dims[a],) + (1,) * (p - 1 - a))
            tw = np.exp(-2j * np.pi * ((ka * mrest) % Q) / Q)
            z = z * tw.reshape((1,) * a + tw.shape)
    return z.reshape(-1)


def inverse_packed(v, dims):
    """Inverse of the digit-transposed complex transform (unnormalised)."""
    p = len(dims)
    z = v.reshape(dims)
    for a in range(p - 1, -1, -1):
        if a < p - 1:
            Q = int(np.prod(dims[a:]))
            mrest = np.arange(Q // dims[a]).reshape(dims[a + 1:])
            ka = np.arange(dims[a]).reshape((dims[a],) + (1,) * (p - 1 - a))
            tw = np.exp(+2j * np.pi * ((ka * mrest) % Q) / Q)
            z = z * tw.reshape((1,) * a + tw.shape)
        z = np.fft.ifft(z, axis=a) * dims[a]
    return z.reshape(-1)


def freq_of_address(dims):
    """k(addr) for the digit-transposed layout."""
    p = len(dims)
    idx = np.indices(dims).reshape(p, -1)
    k = np.zeros(idx.shape[1], dtype=np.int64)
    base = 1
    for a in range(p):
        k += idx[a] * base
        base *= dims[a]
    return k


def mirror_address(addr, dims):
    """Address holding frequency (M - k) mod M, via the region rule
    a' = S_a0 + S_{a0+1} - 1 - a for a in [S_{a0+1}, S_a0)."""
    p = len(dims)
    S = [int(np.prod(dims[a:])) for a in range(p)] + [1]
    addr = np.asarray(addr)
    out = np.zeros_like(addr)
    for a0 in range(p):
        sel = (addr >= S[a0 + 1]) & (addr < S[a0])
        out[sel] = S[a0] + S[a0 + 1] - 1 - addr[sel]
    return out  # addr 0 -> 0


def half_spectrum(Z, dims):
    """Packed transform Z (digit-transposed) -> half spectrum X of the real
    signal in the same address order: address of k holds X[k] = rfft(x)[k],
    k = 0..M-1; address 0 holds X[0] + i X[M] (both real).  This is what the
    forward transform stores (RS_HALF)."""
    M = len(Z)
    a = np.arange(M)
    am = mirror_address(a, dims)
    k = freq_of_address(dims)
    W = np.exp(-2j * np.pi * k / (2 * M))    # W_L^k, L = 2M
    Zc = np.conj(Z[am])
    S, D = Z + Zc, Z - Zc
    X = 0.5 * (S - 1j * W * D)
    X[0] = (Z[0].real + Z[0].imag) + 1j * (Z[0].real - Z[0].imag)
    return X


def pair_product(Xi, Xj, dims, scale=1.0):
    """V (digit-transposed) whose inverse packed transform is the packed
    cross-correlation r[2m] + i r[2m+1] (times M), from the two half spectra."""
    M = len(Xi)
    a = np.arange(M)
    am = mirror_address(a, dims)
    k = freq_of_address(dims)
    C = scale * Xi * np.conj(Xj)
    C0, CM = scale * Xi[0].real * Xj[0].real, scale * Xi[0].imag * Xj[0].imag
    Cm = np.conj(C[am])
    E = 0.5 * (C + Cm)
    O = 0.5 * np.exp(+2j * np.pi * k / (2 * M)) * (C - Cm)
    V = E + 1j * O
    V[0] = 0.5 * (C0 + CM) + 0.5j * (C0 - CM)
    return V


def cc_model(x, y, dims, normalize=True):
    n = len(x)
    M = int(np.prod(dims))
    L = 2 * M
    Zi = half_spectrum(forward_packed(x, dims), dims)
    Zj = half_spectrum(forward_packed(y, dims), dims)
    scale = 1.0 / (np.linalg.norm(x) * np.linalg.norm(y)) if normalize else 1.0
    V = pair_product(Zi, Zj, dims, scale)
    zr = inverse_packed(V, dims) / M
    r = np.empty(L)
    r[0::2] = zr.real
    r[1::2] = zr.imag
    return np.concatenate((r[L - (n - 1):], r[:n]))
