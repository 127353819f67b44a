// sdrb_finish.cuh -- k_finish: everything after the block front end, one warp per (chunk, row).
//
//   1. chunk boundaries: IQ-corrected, NCO-shifted head and end-window samples; sosfiltfilt's odd
//      extension and zi start-up (head), the tail extension (end), the boundary vector zeta
//   2. carries of the 8 forward / 8 anticausal modal states across the tiles of the chunk
//   3. decimated outputs y[k] = partial output of the front end + carry-in response (+ boundary)
//   4. demodulation: fm (pair phase, 2x trigonometric interpolation by a real-input FFT of half
//      the length and its inverse) | am | re | im
//   5. output low-pass (zero initial state per chunk) in 32 lane segments, chained by a
//      Kogge-Stone scan of the segment maps; framing (native or big-endian doubles)
//
// It replaces k_fixup + k_demod for the common shapes (whole tiles, M a power of two <= 1024, at
// most 2 output sections); those two kernels remain the general path.  Nothing leaves the SM
// between the stages: y, the phase row and the filtered row live in the warp's shared memory.
//
// Reference behaviour reproduced: scipy.signal.decimate's sosfiltfilt boundary handling
// (src/dsp/dsp_processor.py:147), src/dsp/demodulation.py:25-68, dsp_processor.py:32-36,149,162,
// src/dsp/vfo_processor.py:84, src/misc/read_file.py:65-77.  tests/emulator.py (emu_fixup,
// emu_demod, emu_fm_interp_real) is the numpy twin.
#pragma once
#include "sdrb_kernels.cuh"

#define FIN_WARPS 4

// Per-warp shared memory (bytes) and per-CTA twiddle table.
__host__ __device__ inline size_t finish_warp_bytes(int M)
{
    const int nt = M / SDRB_TB;
    size_t b = (size_t)(nt + 1) * 16 * sizeof(double2);      // carry
    b += (size_t)M * sizeof(double2);                        // ybuf (addv aliases its head)
    b += 4 * 32 * sizeof(double2);                           // sh, se, seqA, seqB
    b += 8 * sizeof(double2);                                // zeta
    b += 2 * (size_t)(M / 4) * sizeof(double2);              // fa, fb
    b += (size_t)(M + 32) * sizeof(double);                  // zrow, one pad per segment
    return b;
}
__host__ __device__ inline size_t finish_smem_bytes(int M)
{
    return (size_t)(M / 2) * sizeof(double2) + FIN_WARPS * finish_warp_bytes(M);
}

// Stockham radix-2 pass over n points by one warp; twiddle exp(-+2 pi i k / (2 Ns)) = tw[k * tstride].
__device__ __forceinline__ void fin_fft_pass(const double2 *in, double2 *out, int n, int Ns, int tstride,
                                             bool inverse, const double2 *tw, int lane)
{
    const int half = n >> 1;
    for (int j = lane; j < half; j += 32) {
        const int k = j & (Ns - 1);
        double2 w = tw[k * tstride];
        if (inverse) w.y = -w.y;
        const double2 a = in[j], b = cmul(w, in[j + half]);
        const int j0 = ((j - k) << 1) + k;
        out[j0] = cadd(a, b);
        out[j0 + Ns] = csub(a, b);
    }
}

template <int ENC>
__global__ void __launch_bounds__(32 * FIN_WARPS, 2)
k_finish(const __grid_constant__ DevPlan pl, Scratch sc, const uint8_t *__restrict__ raw, double *__restrict__ out,
         int nchunks, int keep_y)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int M = pl.M, nt = pl.ntiles, edge = pl.edge, R = pl.R, h = M >> 1, n2 = M >> 2;
    const bool iq = pl.correct_iq != 0;

    double2 *tw = reinterpret_cast<double2 *>(smem_raw);                 // exp(-2 pi i k / M), k < M/2
    for (int k = threadIdx.x; k < h; k += blockDim.x) tw[k] = pl.tw[k];
    unsigned char *wb = smem_raw + (size_t)h * sizeof(double2) + (size_t)warp * finish_warp_bytes(M);
    double2 *carry = reinterpret_cast<double2 *>(wb);
    double2 *ybuf = carry + (size_t)(nt + 1) * 16;
    double2 *addv = ybuf;                                                // dead before ybuf is written
    double2 *s_h = ybuf + M, *s_e = s_h + 32, *seqA = s_e + 32, *seqB = seqA + 32;
    double2 *s_zeta = seqB + 32;
    double2 *fa = s_zeta + 8, *fb = fa + n2;
    double *zrow = reinterpret_cast<double *>(fb + n2);
    __syncthreads();

    // lane-invariant tables: this lane's block position l = lane inside every tile
    double2 rw[8], rt[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        rw[i] = pl.RW[(size_t)lane * 8 + i];
        rt[i] = pl.RT[(size_t)(SDRB_TB - lane) * 8 + i];
    }
    const int i8 = lane & 7, grp = (lane >> 3) & 1;
    const double2 p_i = pl.p[i8], P32 = pl.Ppow[(size_t)SDRB_TB * 8 + i8];
    const int Ls = pl.sos_Lseg;

    const int gw = blockIdx.x * FIN_WARPS + warp, nw = gridDim.x * FIN_WARPS;
    for (int item = gw; item < nchunks * R; item += nw) {
        const int chunk = item / R, r = item - chunk * R;
        const uint8_t *rawc = raw + (size_t)chunk * pl.N * pl.sb;
        const double2 *offt = sc.off_tile + (size_t)chunk * (nt + 1);
        const double2 *aggr = sc.agg + ((size_t)chunk * R + r) * nt * 16;
        const double2 *T1r = pl.T1 + (size_t)r * nt;

        // ---------------------------------------------------------------- 1a. loads
        if (lane <= edge) {
            s_h[lane] = decode_sample<ENC>(pl, rawc, lane);
            s_e[lane] = decode_sample<ENC>(pl, rawc, (long)pl.ws + lane);
        }
        for (int idx = lane; idx < nt * 16; idx += 32) {
            const int t = idx >> 4, m = idx & 15;
            double2 a = aggr[idx];
            if (iq) {
                const int kind = (t == nt - 1) ? 1 : 0;
                const double2 o = offt[t];
                const double2 psi = m < 8 ? pl.PsiW[((size_t)kind * R + r) * 8 + m]
                                          : pl.PsiT[((size_t)kind * R + r) * 8 + m - 8];
                a = cfma(make_double2(-o.x, -o.y), psi, a);
            }
            addv[idx] = cmul(T1r[t], a);
        }
        __syncwarp();
        // ---------------------------------------------------------------- 1b. IQ recurrences
        // lane 0: head forward from the chunk-start offset (x = z - o; o += L x); lane 1: end
        // window backward from the offset at q*Mf (o = (o - L z) / lam; x = z - o)
        if (iq && lane < 2) {
            const bool fwd = lane == 0;
            double2 *buf = fwd ? s_h : s_e;
            double2 o = fwd ? offt[0] : offt[nt];
            for (int step = 0; step <= edge; step++) {
                const int idx = fwd ? step : edge - step;
                const double2 z = buf[idx];
                double2 x;
                if (fwd) {
                    x = csub(z, o);
                    o.x = fma(x.x, pl.Liq, o.x); o.y = fma(x.y, pl.Liq, o.y);
                } else {
                    o.x = fma(-pl.Liq, z.x, o.x) * pl.lam_inv; o.y = fma(-pl.Liq, z.y, o.y) * pl.lam_inv;
                    x = csub(z, o);
                }
                buf[idx] = x;
            }
        }
        __syncwarp();
        // ---------------------------------------------------------------- 1c. NCO, extensions
        if (lane <= edge) {
            s_h[lane] = cmul(s_h[lane], pl.Ehead[(size_t)r * (edge + 1) + lane]);
            s_e[lane] = cmul(s_e[lane], pl.Eend[(size_t)r * pl.nend + lane]);
        }
        __syncwarp();
        const double2 x0 = s_h[0], xN1 = s_e[edge];
        if (lane < edge) {
            seqA[lane] = csub(cscale(2.0, x0), s_h[edge - lane]);           // odd extension, head
            seqB[lane] = csub(cscale(2.0, xN1), s_e[edge - 1 - lane]);      // odd extension, tail
        }
        __syncwarp();
        // ---------------------------------------------------------------- 1d. head / tail states
        // lanes 0..7: forward modal state at n = edge (zi start-up, then the head extension);
        // lanes 8..15: anticausal state at n = edge + q*Mf from the tail extension
        double2 st = grp ? make_double2(0.0, 0.0) : cmul(pl.zhat[i8], seqA[0]);
        for (int j = 0; j < edge; j++) {
            const double2 v = grp ? seqB[edge - 1 - j] : seqA[j];
            st = cfma(p_i, st, v);
        }
        // ---------------------------------------------------------------- 1e. carries across tiles
        if (lane < 16 && grp) carry[(size_t)nt * 16 + 8 + i8] = st;
        for (int tt = 0; tt < nt; tt++) {
            const int t = grp ? nt - 1 - tt : tt;
            if (lane < 8) carry[(size_t)t * 16 + i8] = st;
            st = cfma(P32, st, addv[t * 16 + grp * 8 + i8]);
            if (lane >= 8 && lane < 16) carry[(size_t)t * 16 + 8 + i8] = st;
        }
        if (lane < 8) carry[(size_t)nt * 16 + i8] = st;
        __syncwarp();
        // ---------------------------------------------------------------- 1f. end, boundary vector
        {
            double2 w = carry[(size_t)nt * 16 + i8], wL1 = w, last = make_double2(0.0, 0.0);
            for (int k = 0; k < edge; k++) {
                const double2 v = seqB[k];
                if (k == edge - 1) { wL1 = w; last = v; }
                w = cfma(p_i, w, v);
            }
            double2 part = cmul(pl.c[i8], wL1);
            for (int sft = 1; sft < 8; sft <<= 1) part = cadd(part, shfl_xor_c(part, sft));
            const double2 yfL1 = make_double2(fma(pl.d, last.x, part.x), fma(pl.d, last.y, part.y));
            double2 zeta = cmul(pl.zhat[i8], yfL1);
#pragma unroll
            for (int l = 0; l < 8; l++) {
                const double2 wl = shfl_c(w, l);
                zeta = csub(zeta, cmul(pl.xi[i8 * 8 + l], wl));
            }
            if (lane < 8) s_zeta[i8] = zeta;
        }
        __syncwarp();
        // ---------------------------------------------------------------- 2. outputs, lane <-> block
        {
            const double2 *ypr = sc.ypart + ((size_t)chunk * R + r) * pl.Mf;
            double2 *yg = sc.y + ((size_t)chunk * R + r) * M;
            const double2 psi0 = pl.psiY[((size_t)0 * R + r) * SDRB_TB + lane];
            const double2 psi1 = pl.psiY[((size_t)1 * R + r) * SDRB_TB + lane];
            for (int t = 0; t < nt; t++) {
                const int k = t * SDRB_TB + lane;
                double2 v = ypr[k];
                if (iq) {
                    const double2 o = offt[t];
                    v = cfma(make_double2(-o.x, -o.y), (t == nt - 1) ? psi1 : psi0, v);
                }
                v = cmul(T1r[t], v);
                const double2 *Win = carry + (size_t)t * 16, *Tn = carry + (size_t)(t + 1) * 16 + 8;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    v = cfma(rw[i], Win[i], v);
                    v = cfma(rt[i], Tn[i], v);
                }
                if (k >= pl.k_bnd) {
#pragma unroll
                    for (int i = 0; i < 8; i++) v = cfma(pl.bnd[(size_t)k * 8 + i], s_zeta[i], v);
                }
                ybuf[k] = v;
                if (keep_y) yg[k] = v;
            }
        }
        __syncwarp();
        // ---------------------------------------------------------------- 3. demodulation
        if (pl.demod == 0) {
            // pair phases (demodulation.py:25-32); even outputs of the 2x interpolation are the
            // phases themselves, the odd ones come from the half-sample-shifted spectrum
            double *ph = reinterpret_cast<double *>(fa);
            for (int i = lane; i < h; i += 32) {
                const double2 a = ybuf[2 * i], b = ybuf[2 * i + 1];
                const double re = fma(a.x, b.x, a.y * b.y), im = fma(a.y, b.x, -a.x * b.y);
                const double v = atan2(im, re);
                ph[i] = v;
                const int ko = 2 * i;
                zrow[ko + ko / Ls] = v;
            }
            __syncwarp();
            // forward FFT of z[m] = ph[2m] + i ph[2m+1], n2 = h/2 points
            double2 *src = fa, *dst = fb;
            for (int Ns = 1; Ns < n2; Ns <<= 1) {
                fin_fft_pass(src, dst, n2, Ns, M / (2 * Ns), false, tw, lane);
                __syncwarp();
                double2 *tmp = src; src = dst; dst = tmp;
            }
            // spectrum of the real row X[k] (k <= h/2) from Z, times the half-sample shift
            // H[k] = exp(i pi k / h), re-packed for the half-length inverse transform
            for (int k = lane; k <= (n2 >> 1); k += 32) {
                if (k == 0) {
                    const double2 Z0 = src[0];
                    const double X0 = Z0.x + Z0.y;
                    dst[0] = make_double2(X0, X0);
                } else {
                    const int kp = n2 - k;
                    const double2 Zk = src[k], Zc = cconj(src[kp]);
                    const double2 Xe = cscale(0.5, cadd(Zk, Zc));
                    const double2 dd = csub(Zk, Zc);
                    const double2 Xo = make_double2(0.5 * dd.y, -0.5 * dd.x);           // -i (Zk - Zc) / 2
                    const double2 w2 = tw[2 * k];                                       // exp(-2 pi i k / h)
                    const double2 tX = cmul(w2, Xo);
                    const double2 Xk = cadd(Xe, tX), Xkp = cconj(csub(Xe, tX));
                    const double2 Uk = cmul(Xk, cconj(tw[k])), Ukp = cmul(Xkp, cconj(tw[kp]));
                    // V[k] = (U[k] + conj U[k']) + i (U[k] - conj U[k']) conj(w2)
                    const double2 s1 = cadd(Uk, cconj(Ukp)), d1 = cmul(csub(Uk, cconj(Ukp)), cconj(w2));
                    dst[k] = make_double2(s1.x - d1.y, s1.y + d1.x);
                    // V[k'] = (U[k'] + conj U[k]) + i (U[k'] - conj U[k]) (-w2)
                    const double2 s2 = cadd(Ukp, cconj(Uk)), d2 = cmul(csub(Ukp, cconj(Uk)), w2);
                    dst[kp] = make_double2(s2.x + d2.y, s2.y - d2.x);
                }
            }
            __syncwarp();
            { double2 *tmp = src; src = dst; dst = tmp; }
            for (int Ns = 1; Ns < n2; Ns <<= 1) {
                fin_fft_pass(src, dst, n2, Ns, M / (2 * Ns), true, tw, lane);
                __syncwarp();
                double2 *tmp = src; src = dst; dst = tmp;
            }
            const double sc1 = 1.0 / (double)h;
            for (int m = lane; m < n2; m += 32) {
                const double2 v = src[m];
                const int k1 = 4 * m + 1, k3 = 4 * m + 3;
                zrow[k1 + k1 / Ls] = v.x * sc1;
                zrow[k3 + k3 / Ls] = v.y * sc1;
            }
        } else {
            for (int k = lane; k < M; k += 32) {
                const double2 a = ybuf[k];
                double v;
                if (pl.demod == 1) v = hypot(fma(a.x, a.x, -a.y * a.y), 2.0 * a.x * a.y);   // abs(square(z))
                else v = pl.demod == 2 ? a.x : a.y;
                zrow[k + k / Ls] = v;
            }
        }
        __syncwarp();
        // ---------------------------------------------------------------- 4. output low-pass
        if (pl.nsec_out > 0) {
            const int nsec = pl.nsec_out;
            double *zs = zrow + (size_t)lane * (Ls + 1);
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            for (int k = 0; k < Ls; k++) {
                double xc = zs[k];
                {
                    const double *cf = pl.out_sos;
                    const double xn = __dadd_rn(__dmul_rn(cf[0], xc), s0);
                    s0 = __dadd_rn(__dadd_rn(__dmul_rn(cf[1], xc), -__dmul_rn(cf[4], xn)), s1);
                    s1 = __dadd_rn(__dmul_rn(cf[2], xc), -__dmul_rn(cf[5], xn));
                    xc = xn;
                }
                if (nsec > 1) {
                    const double *cf = pl.out_sos + 6;
                    const double xn = __dadd_rn(__dmul_rn(cf[0], xc), s2);
                    s2 = __dadd_rn(__dadd_rn(__dmul_rn(cf[1], xc), -__dmul_rn(cf[4], xn)), s3);
                    s3 = __dadd_rn(__dmul_rn(cf[2], xc), -__dmul_rn(cf[5], xn));
                    xc = xn;
                }
                zs[k] = xc;
            }
            // inclusive Kogge-Stone scan of the segment maps s -> A^Ls s + b_lane
#pragma unroll
            for (int lv = 0; lv < 5; lv++) {
                const double o0 = __shfl_up_sync(0xffffffffu, s0, 1 << lv), o1 = __shfl_up_sync(0xffffffffu, s1, 1 << lv);
                const double o2 = __shfl_up_sync(0xffffffffu, s2, 1 << lv), o3 = __shfl_up_sync(0xffffffffu, s3, 1 << lv);
                if (lane >= (1 << lv)) {
                    const double *A = pl.sos_AP + lv * 16;
                    s0 = fma(A[0], o0, fma(A[1], o1, fma(A[2], o2, fma(A[3], o3, s0))));
                    s1 = fma(A[4], o0, fma(A[5], o1, fma(A[6], o2, fma(A[7], o3, s1))));
                    s2 = fma(A[8], o0, fma(A[9], o1, fma(A[10], o2, fma(A[11], o3, s2))));
                    s3 = fma(A[12], o0, fma(A[13], o1, fma(A[14], o2, fma(A[15], o3, s3))));
                }
            }
            double i0 = __shfl_up_sync(0xffffffffu, s0, 1), i1 = __shfl_up_sync(0xffffffffu, s1, 1);
            double i2 = __shfl_up_sync(0xffffffffu, s2, 1), i3 = __shfl_up_sync(0xffffffffu, s3, 1);
            if (lane == 0) { i0 = 0; i1 = 0; i2 = 0; i3 = 0; }
            for (int k = 0; k < Ls; k++) {
                const double *ca = pl.sos_CA + (size_t)k * 4;
                zs[k] = fma(ca[0], i0, fma(ca[1], i1, fma(ca[2], i2, fma(ca[3], i3, zs[k]))));
            }
        }
        __syncwarp();
        // ---------------------------------------------------------------- 5. framing
        double *o = out + ((size_t)r * nchunks + chunk) * M;
        for (int k = lane; k < M; k += 32) {
            const double v = zrow[k + k / Ls];
            o[k] = pl.be_out ? bswap_double(v) : v;
        }
        __syncwarp();
    }
}
