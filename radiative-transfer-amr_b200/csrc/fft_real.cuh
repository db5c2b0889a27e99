// Register-resident real FFTs of length NC = 4 .. 64 (powers of two) for the
// block-circulant angular scattering product (uniform angular meshes).
//
// One thread transforms one sequence x[0..NC) held in registers.  All loops are
// fully unrolled, every array index and twiddle exponent is a compile-time
// constant, so the twiddles become constant-bank operands of DFMA/DMUL and no
// local memory is used (checked with -Xptxas -v: 0 bytes stack).
//
// Conventions (rt/tables.py:circulant_spectrum folds the factors into the table):
//   forward : slot 0 = (X[0], X[NC/2])  -- both real;
//             slot k = 2 X[k], 0 < k < NC/2,  X[k] = sum_K x[K] exp(-2 pi i k K / NC)
//   inverse : x[K] * NC from the true spectrum stored in the same slot layout.
// Method: pack z[m] = x[2m] + i x[2m+1], radix-2 DIT complex FFT of size NC/2,
// then the real/Hermitian split.
#pragma once

// cos / sin (2 pi t / 64), t = 0 .. 31; filled by rt_fft_setup()
__constant__ double c_fft_cos[32];
__constant__ double c_fft_sin[32];

namespace rtfft {

__host__ __device__ constexpr int log2c(int n) { return n <= 1 ? 0 : 1 + log2c(n / 2); }
__host__ __device__ constexpr int bitrev(int i, int bits) {
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
    return r;
}

// in-place radix-2 DIT on bit-reversed input; INV selects exp(+i...) twiddles
template <int M, bool INV>
__device__ __forceinline__ void cfft(double (&zr)[M], double (&zi)[M]) {
#pragma unroll
    for (int half = 1; half < M; half <<= 1) {
#pragma unroll
        for (int g = 0; g < M; g += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const int t = j * (M / (2 * half));          // twiddle W_M^t
                const int i0 = g + j, i1 = g + j + half;
                double br, bi;
                if (t == 0) {
                    br = zr[i1];
                    bi = zi[i1];
                } else if (4 * t == M) {                     // -i (forward), +i (inverse)
                    br = INV ? -zi[i1] : zi[i1];
                    bi = INV ? zr[i1] : -zr[i1];
                } else {
                    const double c = c_fft_cos[t * (64 / M)], s = c_fft_sin[t * (64 / M)];
                    if (!INV) {
                        br = fma(zr[i1], c, zi[i1] * s);
                        bi = fma(zi[i1], c, -(zr[i1] * s));
                    } else {
                        br = fma(zr[i1], c, -(zi[i1] * s));
                        bi = fma(zi[i1], c, zr[i1] * s);
                    }
                }
                const double ar = zr[i0], ai = zi[i0];
                zr[i0] = ar + br;
                zi[i0] = ai + bi;
                zr[i1] = ar - br;
                zi[i1] = ai - bi;
            }
        }
    }
}

template <int NC>
struct Real {
    static constexpr int M = NC / 2;
    static constexpr int BITS = log2c(M);

    __device__ __forceinline__ static void forward(const double (&x)[NC], double (&re)[M],
                                                   double (&im)[M]) {
        double zr[M], zi[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            zr[bitrev(m, BITS)] = x[2 * m];
            zi[bitrev(m, BITS)] = x[2 * m + 1];
        }
        cfft<M, false>(zr, zi);
        re[0] = zr[0] + zi[0];
        im[0] = zr[0] - zi[0];
#pragma unroll
        for (int k = 1; k <= M / 2; ++k) {
            const int mk = M - k;
            const double ar = zr[k] + zr[mk], ai = zi[k] - zi[mk];
            const double br = zr[k] - zr[mk], bi = zi[k] + zi[mk];
            const double c = c_fft_cos[k * (64 / NC)], s = c_fft_sin[k * (64 / NC)];
            const double tr = fma(c, bi, -(s * br));
            const double ti = -fma(c, br, s * bi);
            re[k] = ar + tr;
            im[k] = ai + ti;
            if (mk != k) {
                re[mk] = ar - tr;
                im[mk] = ti - ai;
            }
        }
    }

    __device__ __forceinline__ static void inverse(const double (&re)[M], const double (&im)[M],
                                                   double (&x)[NC]) {
        double zr[M], zi[M];
        zr[0] = re[0] + im[0];
        zi[0] = re[0] - im[0];
#pragma unroll
        for (int k = 1; k <= M / 2; ++k) {
            const int mk = M - k;
            const double pr = re[k] + re[mk], pi = im[k] - im[mk];
            const double qr = re[k] - re[mk], qi = im[k] + im[mk];
            const double c = c_fft_cos[k * (64 / NC)], s = c_fft_sin[k * (64 / NC)];
            const double rr = -fma(c, qi, s * qr);
            const double ri = fma(c, qr, -(s * qi));
            zr[bitrev(k, BITS)] = pr + rr;
            zi[bitrev(k, BITS)] = pi + ri;
            if (mk != k) {
                zr[bitrev(mk, BITS)] = pr - rr;
                zi[bitrev(mk, BITS)] = ri - pi;
            }
        }
        // zr[0] / zi[0] are already in place (bitrev(0) = 0)
        cfft<M, true>(zr, zi);
#pragma unroll
        for (int m = 0; m < M; ++m) {
            x[2 * m] = zr[m];
            x[2 * m + 1] = zi[m];
        }
    }
};

}  // namespace rtfft
