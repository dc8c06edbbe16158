"""The index arithmetic of the paired / summed four-step FFT (csrc/mcd_large.cuh: fft4p_cols_fwd_kernel,
fft4p_rows_kernel, fft4_cols_inv_kernel) stated in NumPy and checked against per-chain circular autocorrelations:

* chains 2q and 2q+1 ride one complex transform z = a + i b and |A(k)|^2 + |B(k)|^2 = (|Z(k)|^2 + |Z(N-k)|^2) / 2;
* with k = k1 + N1 k2 the mirror frequency N - k of row k1 > 0 sits in row N1 - k1 at k2' = N2 - 1 - k2, row 0 mirrors
  onto itself at k2' = (N2 - k2) mod N2;
* the summed power spectrum is inverted once: rows (times W^(-n2 k1)), then columns, n = N2 n1 + n2;
* lags <= maxlag are free of wrap-around as soon as N >= niter + maxlag (the reference pads to 2 niter - 1,
  src/ess_rhat.jl:103-118, for lags it never reads, :181-195);
* mean_i(c[k,i] / c[0,i] var_i) = (sum_i c[k,i] / sum_i c[0,i]) mean_i(var_i) because c[0,i] = (niter - 1) var_i.
CPU only: this pins the math the kernels implement; the kernels themselves are checked on the GPU
(tests/test_gpu_fft_paths.py)."""
import numpy as np
import pytest


def four_step_paired(Y, N1, N2, maxlag):
    nch, niter = Y.shape
    N = N1 * N2
    W = lambda m: np.exp(-2j * np.pi * m / N)
    k1 = np.arange(N1)[:, None]; n2 = np.arange(N2)[None, :]
    P = np.zeros((N1, N2))
    for q in range((nch + 1) // 2):
        a = np.zeros(N); a[:niter] = Y[2 * q]
        b = np.zeros(N)
        if 2 * q + 1 < nch:
            b[:niter] = Y[2 * q + 1]
        A = (a + 1j * b).reshape(N1, N2)                    # n = N2 n1 + n2
        B = np.fft.fft(A, axis=0) * W(n2 * k1)               # columns, twiddle
        Z = np.fft.fft(B, axis=1)                            # rows: Z[k1][k2] = Z(k1 + N1 k2)
        for r in range(N1 // 2 + 1):                         # a CTA: row r and its mirror row
            rm = (N1 - r) % N1
            k2 = np.arange(N2)
            m2 = (N2 - k2) % N2 if r == 0 else N2 - 1 - k2
            p = 0.5 * (np.abs(Z[r]) ** 2 + np.abs(Z[rm][m2]) ** 2)
            P[r] += p
            if rm != r:
                P[rm] += p[N2 - 1 - k2]                      # the power of the mirror row is the row's, read backwards
    R = np.fft.ifft(P, axis=1) * N2 * np.conj(W(n2 * k1))    # inverse rows, twiddle
    C = np.fft.ifft(R, axis=0) * N1                          # inverse columns
    return (C.reshape(-1) / N).real[: maxlag + 1]


@pytest.mark.parametrize("N1,N2,niter,nch,maxlag", [(8, 16, 70, 5, 20), (6, 12, 50, 4, 22), (9, 8, 40, 3, 30), (16, 16, 200, 8, 56),
                                                      (4, 27, 80, 1, 28), (3, 32, 60, 2, 36)])
def test_paired_four_step_equals_per_chain_autocorrelation(N1, N2, niter, nch, maxlag):
    assert N1 * N2 >= niter + maxlag
    rng = np.random.default_rng(N1 * 100 + N2)
    Y = rng.standard_normal((nch, niter)); Y -= Y.mean(axis=1, keepdims=True)
    ref = np.zeros(maxlag + 1)
    for j in range(nch):                                     # linear autocorrelation, no padding tricks
        ref += np.array([np.dot(Y[j, : niter - k], Y[j, k:]) for k in range(maxlag + 1)])
    got = four_step_paired(Y, N1, N2, maxlag)
    assert np.allclose(got, ref, rtol=1e-10, atol=1e-10)


def test_reference_weighting_equals_the_summed_form():
    rng = np.random.default_rng(3)
    niter, nch, maxlag = 300, 6, 40
    Y = rng.standard_normal((nch, niter)) * rng.uniform(0.5, 2.0, (nch, 1)); Y -= Y.mean(axis=1, keepdims=True)
    var = Y.var(axis=1, ddof=1)
    c = np.array([[np.dot(Y[j, : niter - k], Y[j, k:]) for k in range(maxlag + 1)] for j in range(nch)])
    reference = np.mean(c / c[:, :1] * var[:, None], axis=0) * (niter - 1) / niter       # src/ess_rhat.jl:181-195
    summed = c.sum(axis=0) / c[:, 0].sum() * var.mean() * (niter - 1) / niter
    assert np.allclose(reference, summed, rtol=1e-12)
