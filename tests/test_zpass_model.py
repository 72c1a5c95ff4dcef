"""CPU model of the hand-written z leg of the Poisson solve (csrc/poisson_z.cu): the Stockham stage / radix
plan, the index formulas a thread uses, the conj-FFT-conj inverse with reversed radices, and the shared-memory
row mapping.  The kernel itself is checked against cuFFT on the GPU (test_fused_z_pass_matches_cufft_z_leg);
this file pins the arithmetic it is built from, without a GPU."""
import numpy as np
import pytest

PLANS = {4: [4, 4], 5: [8, 4], 6: [8, 8], 7: [8, 4, 4], 8: [8, 8, 4], 9: [8, 8, 8], 10: [8, 8, 4, 4]}  # Plan<LOGN>


def dft_small(v):
    r = len(v)
    k = np.arange(r)
    return np.array([np.sum(v * np.exp(-2j * np.pi * k * q / r)) for q in range(r)])


def stockham(x, radices):
    """One transform exactly as run_stages() walks it: thread t of N/8, 8/R butterflies per radix-R stage."""
    n = len(x)
    tw = np.exp(-2j * np.pi * np.arange(n) / n)  # k_twiddles
    data, ns = x.copy(), 1
    for r in radices:
        out = np.zeros_like(data)
        for t in range(n // 8):
            for u in range(8 // r):
                j = t + u * (n // 8)
                v = np.array([data[j + q * (n // r)] for q in range(r)])  # slot_index
                k = j & (ns - 1)
                step = k * (n // (r * ns))
                for q in range(1, r):
                    v[q] *= tw[q * step]  # stage_compute
                v = dft_small(v)
                j0 = (j - k) * r + k
                for q in range(r):
                    out[j0 + q * ns] = v[q]  # stage_store
        data, ns = out, ns * r
    return data


@pytest.mark.parametrize("logn", sorted(PLANS))
def test_stage_plan_is_a_natural_order_fft_and_its_inverse(logn):
    n = 1 << logn
    rng = np.random.default_rng(logn)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    g = rng.standard_normal(n)  # a real influence function
    plan = PLANS[logn]
    assert np.prod(plan) == n
    y = stockham(x, plan)
    assert np.abs(y - np.fft.fft(x)).max() < 1e-10 * n
    # multiply in place, then inverse = conj(FFT(conj(.))) with the radices reversed; the last forward stage
    # (Ns = N / R) leaves thread t holding j + r N/R -- the inputs of the first stage of that inverse
    z = np.conj(stockham(np.conj(y * g), plan[::-1])) / n
    assert np.abs(z - np.fft.ifft(np.fft.fft(x) * g)).max() < 1e-10


def sidx(pos, c, cols):
    return (pos + (pos >> 3)) * cols + c


@pytest.mark.parametrize("logn,cols", [(9, 8), (6, 32), (8, 16), (7, 32), (10, 4)])
def test_shared_memory_rows_are_bank_conflict_free_for_fp32(logn, cols):
    """A warp = `cols` columns x 32/cols consecutive threads t.  With radix-8 first stages in both directions
    (N = 512, 64) every access of the kernel (stage loads, stage stores) puts the 32 word addresses of a warp
    into 32 distinct banks; plans that end in a radix-4 stage (its store opens the inverse with stride 4) and the
    4-column variant (8-row seams) are allowed a 2-way conflict there."""
    n = 1 << logn
    plan = PLANS[logn]
    tpw = 32 // cols
    worst = 1
    for inverse in (False, True):
        radices = plan[::-1] if inverse else plan
        ns = 1
        for s, r in enumerate(radices):
            for t0 in range(0, n // 8, tpw):
                ts = np.arange(t0, t0 + tpw)
                for u in range(8 // r):
                    j = ts + u * (n // 8)
                    for q in range(r):
                        patterns = []
                        if s > 0:  # stage_load
                            patterns.append(j + q * (n // r))
                        if s < len(radices) - 1:  # stage_store
                            k = j & (ns - 1)
                            patterns.append((j - k) * r + k + q * ns)
                        for pos in patterns:
                            banks = [sidx(int(p), c, cols) % 32 for p in pos for c in range(cols)]
                            worst = max(worst, max(np.bincount(banks, minlength=32)))
            ns *= r
    assert worst <= (1 if cols >= 8 and plan[-1] == 8 else 2)
