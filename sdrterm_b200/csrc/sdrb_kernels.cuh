// sdrb_kernels.cuh -- the sm_100a kernels of the demodulation chain.
//
//   k_main    decode + block-local IQ + NCO + even/odd block sums on the FP64 tensor pipe (DMMA)
//             + tile-local modal scans -> partial outputs and tile aggregates
//   k_iqscan  IQ-corrector offset at every tile start (carried across chunks and calls)
//   k_fixup   head/end segments, cross-tile carries, boundary term -> decimated complex y
//   k_demod   fm (pair phase + 2x FFT interpolation) | am | re | im, output SOS, framing
//
// Reference behaviour being reproduced: src/misc/read_file.py:100-103 (decode, normalise, IQ
// correction), src/dsp/demodulation.py:71-79 (NCO), src/dsp/dsp_processor.py:147 (scipy
// decimate), demodulation.py:25-68, dsp_processor.py:32-36,149,162, vfo_processor.py:84.
// tests/emulator.py is the line-by-line numpy twin of these kernels.
#pragma once
#include "sdrb_device.cuh"

// Scratch written by k_main / read by k_fixup, per batch of chunks.
struct Scratch {
    double2 *ypart;     // [nch][R][Mf]
    double2 *agg;       // [nch][R][ntiles][16]   0..7 Wout, 8..15 Tin
    double2 *tile_agg;  // [nch][ntiles]
    double2 *tailwin;   // [nch][edge+1]
    double2 *off_tile;  // [nch][ntiles+1]
    double2 *carry;     // [nch][R][ntiles+1][16] 0..7 Win[t], 8..15 Tn[t]
    double2 *y;         // [nch][R][M]
    double2 *iq_state;  // [1] offset before the batch (in) / after it (out)
    double2 *fftbuf;    // [nch*R][2][M] when the demod buffers do not fit shared memory
    double *zrow;       // [nch*R][M]     "
};

// ------------------------------------------------------------------------------------ k_main
// grid.x = nchunks * ceil(ntiles / TPC); block = 32*W threads.  A CTA decodes TPC consecutive
// tiles of one chunk into shared memory (phase 0, once, shared by all rows) and its warps then
// take (tile, row) items.
template <int ENC>
__global__ void __launch_bounds__(256)
k_main(const __grid_constant__ DevPlan pl, Scratch sc, const uint8_t *__restrict__ raw, int nchunks, int TPC)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int q = pl.q, zs = SDRB_TB + SDRB_ZPAD;
    const int groups = (pl.ntiles + TPC - 1) / TPC;
    const int chunk = blockIdx.x / groups, tg = blockIdx.x % groups;
    if (chunk >= nchunks) return;
    double2 *zT = reinterpret_cast<double2 *>(smem_raw);                 // [TPC][q][zs]
    double2 *cl = zT + (size_t)TPC * q * zs;                              // [TPC][32]
    double *xball = reinterpret_cast<double *>(cl + TPC * SDRB_TB);       // [W][32][XSTRIDE]
    const uint8_t *rawc = raw + (size_t)chunk * pl.N * 2 * pl.itemsize;
    const int t0 = tg * TPC;
    const int ntl = min(TPC, pl.ntiles - t0);

    // ---------------- phase 0: decode + block-local IQ correction, lane <-> block
    for (int tl = warp; tl < ntl; tl += W) {
        const int t = t0 + tl;
        const int cnt = (t == pl.ntiles - 1) ? pl.cnt_last : SDRB_TB;
        const bool active = lane < cnt;
        const long base = ((long)t * SDRB_TB + lane) * q;
        double2 *zTt = zT + (size_t)tl * q * zs;
        double2 acc = make_double2(0.0, 0.0);
        for (int j = 0; j < q; j++) {
            double2 z = make_double2(0.0, 0.0);
            if (active) z = decode_sample<ENC>(pl, rawc, base + j);
            double2 zp = z;
            if (pl.correct_iq) {
                zp.x = fma(-pl.Liq, acc.x, z.x);
                zp.y = fma(-pl.Liq, acc.y, z.y);
                acc.x = fma(pl.lam, acc.x, z.x);
                acc.y = fma(pl.lam, acc.y, z.y);
            }
            zTt[j * zs + lane] = zp;
        }
        double2 excl = make_double2(0.0, 0.0);
        if (pl.correct_iq) {
            double2 inc = active ? cscale(pl.Liq, acc) : make_double2(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < 5; i++) {
                double2 tt = shfl_up_c(inc, 1 << i);
                if (lane >= (1 << i)) { inc.x = fma(pl.lamq_pow[i], tt.x, inc.x); inc.y = fma(pl.lamq_pow[i], tt.y, inc.y); }
            }
            excl = shfl_up_c(inc, 1);
            if (lane == 0) excl = make_double2(0.0, 0.0);
            double2 tagg = shfl_c(inc, cnt - 1);
            if (lane == 0) sc.tile_agg[(size_t)chunk * pl.ntiles + t] = tagg;
            // tail window: tile-local-corrected samples n in [N-1-edge, q*Mf)
            const long n0 = pl.N - 1 - pl.edge;
            if (active && base + q > n0) {
                for (int j = 0; j < q; j++) {
                    long n = base + j;
                    if (n >= n0) {
                        double2 v = zTt[j * zs + lane];
                        double lj = pl.lam_j[j];
                        sc.tailwin[(size_t)chunk * (pl.edge + 1) + (n - n0)] =
                            make_double2(fma(-lj, excl.x, v.x), fma(-lj, excl.y, v.y));
                    }
                }
            }
        } else {
            const long n0 = pl.N - 1 - pl.edge;
            if (active && base + q > n0)
                for (int j = 0; j < q; j++)
                    if (base + j >= n0)
                        sc.tailwin[(size_t)chunk * (pl.edge + 1) + (base + j - n0)] = zTt[j * zs + lane];
        }
        cl[tl * SDRB_TB + lane] = excl;
    }
    __syncthreads();

    // ---------------- items: (tile, row)
    double *xb = xball + (size_t)warp * 32 * SDRB_XSTRIDE;
    const int e = (lane >> 2) & 1, m = lane >> 3;
    for (int item = warp; item < ntl * pl.R; item += W) {
        const int tl = item % ntl, r = item / ntl;
        const int t = t0 + tl;
        const int cnt = (t == pl.ntiles - 1) ? pl.cnt_last : SDRB_TB;
        const double2 *zTt = zT + (size_t)tl * q * zs;
        const double2 *T2r = pl.T2 + (size_t)r * q;
        double acc[4][4][2];
#pragma unroll
        for (int g = 0; g < 4; g++)
#pragma unroll
            for (int ty = 0; ty < 4; ty++) { acc[g][ty][0] = 0.0; acc[g][ty][1] = 0.0; }

        for (int s = 0; s < pl.KS; s++) {
            int j = 4 * s + (lane & 3);
            if (j >= pl.Hq) j = 0;                      // padded pair: coefficient is zero
            const int jm = q - 1 - j;
            const bool mid = (j == jm);
            const double aE = __ldg(pl.Afrag + (size_t)(2 * s) * 32 + lane);
            const double aO = __ldg(pl.Afrag + (size_t)(2 * s + 1) * 32 + lane);
            const double2 t2a = __ldg(T2r + j), t2b = __ldg(T2r + jm);
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const int bB = 8 * g + (lane >> 2);
                const double2 za = zTt[j * zs + bB], zb = zTt[jm * zs + bB];
                const double2 ua = cmul(t2a, za), ub = cmul(t2b, zb);
                double2 a, d;
                if (mid) { a = ua; d = make_double2(0.0, 0.0); }
                else { a = cadd(ua, ub); d = csub(ua, ub); }
                dmma884(acc[g][0][0], acc[g][0][1], aE, a.x);
                dmma884(acc[g][1][0], acc[g][1][1], aE, a.y);
                dmma884(acc[g][2][0], acc[g][2][1], aO, d.x);
                dmma884(acc[g][3][0], acc[g][3][1], aO, d.y);
            }
        }
        // combine the four real sums into F/G (part e of poles m and m+4), local IQ, exchange
        const double2 phFu = pl.PhiF[(size_t)r * 8 + m], phFl = pl.PhiF[(size_t)r * 8 + m + 4];
        const double2 phGu = pl.PhiG[(size_t)r * 8 + m], phGl = pl.PhiG[(size_t)r * 8 + m + 4];
#pragma unroll
        for (int g = 0; g < 4; g++) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int b = 8 * g + 2 * (lane & 3) + i;
                const double Sr = acc[g][0][i], Si = acc[g][1][i], Dr = acc[g][2][i], Di = acc[g][3][i];
                const double oS = __shfl_xor_sync(0xffffffffu, Si, 4);
                const double oD = __shfl_xor_sync(0xffffffffu, Di, 4);
                double Sup, Slo, Dup, Dlo;
                if (e == 0) { Sup = Sr - oS; Slo = Sr + oS; Dup = Dr - oD; Dlo = Dr + oD; }
                else        { Sup = oS + Sr; Slo = oS - Sr; Dup = oD + Dr; Dlo = oD - Dr; }
                double Fu = Sup + Dup, Gu = Sup - Dup, Fl = Slo + Dlo, Gl = Slo - Dlo;
                if (pl.correct_iq) {
                    const double2 cb = cl[tl * SDRB_TB + b];
                    if (e == 0) {
                        Fu -= fma(cb.x, phFu.x, -cb.y * phFu.y); Fl -= fma(cb.x, phFl.x, -cb.y * phFl.y);
                        Gu -= fma(cb.x, phGu.x, -cb.y * phGu.y); Gl -= fma(cb.x, phGl.x, -cb.y * phGl.y);
                    } else {
                        Fu -= fma(cb.x, phFu.y, cb.y * phFu.x); Fl -= fma(cb.x, phFl.y, cb.y * phFl.x);
                        Gu -= fma(cb.x, phGu.y, cb.y * phGu.x); Gl -= fma(cb.x, phGl.y, cb.y * phGl.x);
                    }
                }
                xb[(2 * m + e) * SDRB_XSTRIDE + b] = Fu;
                xb[(2 * (m + 4) + e) * SDRB_XSTRIDE + b] = Fl;
                xb[(2 * (8 + m) + e) * SDRB_XSTRIDE + b] = Gu;
                xb[(2 * (12 + m) + e) * SDRB_XSTRIDE + b] = Gl;
            }
        }
        __syncwarp();
        // tile-local scans in the rotating frame: lanes 0..7 forward poles, 8..15 backward
        if (lane < 16) {
            const bool fwd = lane < 8;
            const double2 Pm = pl.Prot[(size_t)r * 16 + lane];
            double2 st = make_double2(0.0, 0.0);
            double *xr = xb + (2 * lane) * SDRB_XSTRIDE, *xi = xr + SDRB_XSTRIDE;
            for (int step = 0; step < cnt; step++) {
                const int l = fwd ? step : cnt - 1 - step;
                const double2 v = make_double2(xr[l], xi[l]);
                const double2 nst = cfma(Pm, st, v);
                const double2 o = fwd ? st : nst;
                xr[l] = o.x; xi[l] = o.y;
                st = nst;
            }
            if (fwd) st = cmul(pl.T3[(size_t)r * (SDRB_TB + 1) + cnt - 1], st);
            sc.agg[(((size_t)chunk * pl.R + r) * pl.ntiles + t) * 16 + lane] = st;
        }
        __syncwarp();
        // partial outputs, lane <-> block
        if (lane < cnt) {
            double2 sw = make_double2(0.0, 0.0), sT = make_double2(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const double2 Wv = make_double2(xb[(2 * i) * SDRB_XSTRIDE + lane], xb[(2 * i + 1) * SDRB_XSTRIDE + lane]);
                const double2 Tv = make_double2(xb[(2 * (8 + i)) * SDRB_XSTRIDE + lane], xb[(2 * (8 + i) + 1) * SDRB_XSTRIDE + lane]);
                sw = cfma(pl.rho[i], Wv, sw);
                sT = cfma(pl.rho_p[i], Tv, sT);
            }
            const double2 epsb = cconj(pl.T3[(size_t)r * (SDRB_TB + 1) + 1]);
            double2 x0 = csub(zTt[lane], cl[tl * SDRB_TB + lane]);
            double2 ys = cfma(epsb, sw, sT);
            ys.x = fma(pl.g0, x0.x, ys.x); ys.y = fma(pl.g0, x0.y, ys.y);
            const double2 yp = cmul(pl.T3[(size_t)r * (SDRB_TB + 1) + lane], ys);
            sc.ypart[((size_t)chunk * pl.R + r) * pl.Mf + (size_t)t * SDRB_TB + lane] = yp;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------- k_iqscan
// Offset of the IQ corrector at every tile start of every chunk.  One CTA; thread <-> chunk
// group.  State before the batch in sc.iq_state[0]; state after it is written back.
template <int ENC>
__global__ void __launch_bounds__(1024)
k_iqscan(const __grid_constant__ DevPlan pl, Scratch sc, const uint8_t *__restrict__ raw, int nchunks)
{
    __shared__ double2 s_a[1024];   // per-thread aggregate (offset gained from zero over its chunks)
    __shared__ double s_m[1024];    // per-thread decay multiplier
    const int tid = threadIdx.x, nth = blockDim.x;
    const int per = (nchunks + nth - 1) / nth;
    const int c0 = tid * per, c1 = min(nchunks, c0 + per);
    const int nt = pl.ntiles;
    // pass 1: per chunk aggregate (from zero) over tiles + partial block; thread aggregate
    double2 a = make_double2(0.0, 0.0);
    double mlt = 1.0;
    for (int c = c0; c < c1; c++) {
        double2 o = make_double2(0.0, 0.0);
        for (int t = 0; t < nt; t++) {
            const double lt = pl.lam_tile[t == nt - 1 ? 1 : 0];
            const double2 g = sc.tile_agg[(size_t)c * nt + t];
            o.x = fma(lt, o.x, g.x); o.y = fma(lt, o.y, g.y);
        }
        // partial block: reference recurrence from zero gives the affine part; decay lam^rem
        if (pl.rem) {
            const uint8_t *rawc = raw + (size_t)c * pl.N * 2 * pl.itemsize;
            double2 acc = make_double2(0.0, 0.0);
            for (int j = 0; j < pl.rem; j++) {
                double2 z = decode_sample<ENC>(pl, rawc, (long)pl.q * pl.Mf + j);
                acc.x = fma(pl.lam, acc.x, z.x); acc.y = fma(pl.lam, acc.y, z.y);
            }
            const double lr = pl.lam_j[pl.rem];
            o.x = fma(lr, o.x, pl.Liq * acc.x); o.y = fma(lr, o.y, pl.Liq * acc.y);
        }
        // o = offset gained over chunk c from zero; stash it in off_tile[c][nt] for pass 2
        sc.off_tile[(size_t)c * (nt + 1) + nt] = o;
        a.x = fma(pl.lam_N, a.x, o.x); a.y = fma(pl.lam_N, a.y, o.y);
        mlt *= pl.lam_N;
    }
    s_a[tid] = a; s_m[tid] = mlt;
    __syncthreads();
    // serial combine over threads by thread 0 (<= 1024 steps), exclusive prefix into s_a
    if (tid == 0) {
        double2 o = sc.iq_state[0];
        for (int i = 0; i < nth; i++) {
            const double2 ai = s_a[i]; const double mi = s_m[i];
            s_a[i] = o;
            o.x = fma(mi, o.x, ai.x); o.y = fma(mi, o.y, ai.y);
        }
        sc.iq_state[0] = o;
    }
    __syncthreads();
    // pass 2: offsets at tile starts
    double2 o = s_a[tid];
    for (int c = c0; c < c1; c++) {
        const double2 gain = sc.off_tile[(size_t)c * (nt + 1) + nt];
        double2 ot = o;
        for (int t = 0; t < nt; t++) {
            sc.off_tile[(size_t)c * (nt + 1) + t] = ot;
            const double lt = pl.lam_tile[t == nt - 1 ? 1 : 0];
            const double2 g = sc.tile_agg[(size_t)c * nt + t];
            ot.x = fma(lt, ot.x, g.x); ot.y = fma(lt, ot.y, g.y);
        }
        sc.off_tile[(size_t)c * (nt + 1) + nt] = ot;      // offset at sample q*Mf
        o.x = fma(pl.lam_N, o.x, gain.x); o.y = fma(pl.lam_N, o.y, gain.y);
    }
}

// ----------------------------------------------------------------------------------- k_fixup
// One CTA per (chunk, row).  Warp 0: lanes 0..7 <-> poles run the head, the cross-tile carries,
// the end segment and the boundary vector zeta; then all threads emit y[k].
template <int ENC>
__global__ void __launch_bounds__(128)
k_fixup(const __grid_constant__ DevPlan pl, Scratch sc, const uint8_t *__restrict__ raw, int nchunks)
{
    __shared__ double2 s_x[320];       // head samples (edge+1) then end-window samples (nend)
    __shared__ double2 s_zeta[SDRB_NP];
    const int chunk = blockIdx.x / pl.R, r = blockIdx.x % pl.R;
    if (chunk >= nchunks) return;
    const int nt = pl.ntiles, edge = pl.edge, q = pl.q;
    const uint8_t *rawc = raw + (size_t)chunk * pl.N * 2 * pl.itemsize;
    const double2 *offt = sc.off_tile + (size_t)chunk * (nt + 1);
    const double2 *aggr = sc.agg + ((size_t)chunk * pl.R + r) * nt * 16;
    double2 *carry = sc.carry + ((size_t)chunk * pl.R + r) * (nt + 1) * 16;
    double2 *yrow = sc.y + ((size_t)chunk * pl.R + r) * pl.M;
    const double2 *T1r = pl.T1 + (size_t)r * nt;
    double2 *s_h = s_x, *s_e = s_x + (edge + 1);

    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        // corrected + shifted head and end-window samples (lane 0 and lane 1, serial recurrences)
        if (lane == 0) {
            double2 o = pl.correct_iq ? offt[0] : make_double2(0.0, 0.0);
            for (int n = 0; n <= edge; n++) {
                double2 z = decode_sample<ENC>(pl, rawc, n);
                double2 x = csub(z, o);
                o.x = fma(x.x, pl.Liq, o.x); o.y = fma(x.y, pl.Liq, o.y);
                s_h[n] = cmul(x, pl.Ehead[(size_t)r * (edge + 1) + n]);
            }
        } else if (lane == 1) {
            double2 o = pl.correct_iq ? offt[nt] : make_double2(0.0, 0.0);
            const long n0 = pl.N - 1 - edge;
            for (int i = 0; i < pl.nend; i++) {
                const long n = pl.ws + i;
                double2 x;
                if (n < (long)q * pl.Mf) {
                    x = sc.tailwin[(size_t)chunk * (edge + 1) + (n - n0)];
                    if (pl.correct_iq) {
                        const int t = (int)((n / q) / SDRB_TB);
                        const double lp = pow(pl.lam, (double)(n - (long)t * SDRB_TB * q));
                        x.x = fma(-lp, offt[t].x, x.x); x.y = fma(-lp, offt[t].y, x.y);
                    }
                } else {
                    double2 z = decode_sample<ENC>(pl, rawc, n);
                    x = csub(z, o);
                    o.x = fma(x.x, pl.Liq, o.x); o.y = fma(x.y, pl.Liq, o.y);
                }
                s_e[i] = cmul(x, pl.Eend[(size_t)r * pl.nend + i]);
            }
        }
        __syncwarp();
        const int i = lane & 7;                       // pole handled by this lane (lanes >= 8 mirror)
        const double2 p = pl.p[i];
        // head: odd extension, zi, edge samples -> state at n = edge
        const double2 x0 = s_h[0];
        double2 w = cmul(pl.zhat[i], csub(cscale(2.0, x0), s_h[edge]));
        for (int j = 0; j < edge; j++) {
            const double2 ext = csub(cscale(2.0, x0), s_h[edge - j]);
            w = cfma(p, w, ext);
        }
        // forward carries across tiles
        for (int t = 0; t < nt; t++) {
            if (lane < 8) carry[(size_t)t * 16 + i] = w;
            const int kind = (t == nt - 1) ? 1 : 0;
            const int cnt = kind ? pl.cnt_last : SDRB_TB;
            double2 add = aggr[(size_t)t * 16 + i];
            if (pl.correct_iq) {
                const double2 s = make_double2(-offt[t].x, -offt[t].y);
                add = cfma(s, pl.PsiW[((size_t)kind * pl.R + r) * 8 + i], add);
            }
            w = cfma(pl.Ppow[(size_t)cnt * 8 + i], w, cmul(T1r[t], add));
        }
        if (lane < 8) carry[(size_t)nt * 16 + i] = w;
        const double2 wEnd = w;
        // end segment: partial block then tail extension
        const int nseq = pl.rem + edge;
        const double2 xN1 = s_e[pl.nend - 1];
        double2 wL1 = w, last = make_double2(0.0, 0.0);
        for (int k = 0; k < nseq; k++) {
            double2 v = (k < pl.rem) ? s_e[pl.nend - pl.rem + k]
                                     : csub(cscale(2.0, xN1), s_e[pl.nend - 2 - (k - pl.rem)]);
            if (k == nseq - 1) { wL1 = w; last = v; }
            w = cfma(p, w, v);
        }
        const double2 wL = w;
        double2 T = make_double2(0.0, 0.0);
        for (int k = nseq - 1; k >= 0; k--) {
            double2 v = (k < pl.rem) ? s_e[pl.nend - pl.rem + k]
                                     : csub(cscale(2.0, xN1), s_e[pl.nend - 2 - (k - pl.rem)]);
            T = cfma(p, T, v);
        }
        const double2 Tend = T;
        // y_f[L-1] = sum_l c_l w_l[L-1] + d ext[L-1]   (reduce over the 8 pole lanes)
        double2 part = cmul(pl.c[i], wL1);
        for (int sft = 1; sft < 8; sft <<= 1) part = cadd(part, shfl_xor_c(part, sft));
        const double2 yfL1 = make_double2(fma(pl.d, last.x, part.x), fma(pl.d, last.y, part.y));
        double2 zeta = cmul(pl.zhat[i], yfL1);
        for (int l = 0; l < 8; l++) {
            const double2 wl = shfl_c(wL, l);
            const double2 xv = pl.xi[i * 8 + l];
            zeta = csub(zeta, cmul(xv, wl));
        }
        if (lane < 8) s_zeta[i] = zeta;
        if (pl.rem) {
            double2 v = cfma(pl.rho[i], wEnd, cmul(pl.rho_p[i], Tend));
            v = cfma(pl.bnd[(size_t)pl.Mf * 8 + i], zeta, v);
            for (int sft = 1; sft < 8; sft <<= 1) v = cadd(v, shfl_xor_c(v, sft));
            if (lane == 0) {
                const double2 xp = s_e[pl.nend - pl.rem];
                yrow[pl.Mf] = make_double2(fma(pl.g0, xp.x, v.x), fma(pl.g0, xp.y, v.y));
            }
        }
        // backward carries
        T = Tend;
        if (lane < 8) carry[(size_t)nt * 16 + 8 + i] = T;
        for (int t = nt - 1; t >= 0; t--) {
            const int kind = (t == nt - 1) ? 1 : 0;
            const int cnt = kind ? pl.cnt_last : SDRB_TB;
            double2 add = aggr[(size_t)t * 16 + 8 + i];
            if (pl.correct_iq) {
                const double2 s = make_double2(-offt[t].x, -offt[t].y);
                add = cfma(s, pl.PsiT[((size_t)kind * pl.R + r) * 8 + i], add);
            }
            T = cfma(pl.Ppow[(size_t)cnt * 8 + i], T, cmul(T1r[t], add));
            if (lane < 8) carry[(size_t)t * 16 + 8 + i] = T;
        }
    }
    __syncthreads();
    // outputs at the full-block starts
    const double2 *ypr = sc.ypart + ((size_t)chunk * pl.R + r) * pl.Mf;
    for (int k = threadIdx.x; k < pl.Mf; k += blockDim.x) {
        const int t = k / SDRB_TB, l = k % SDRB_TB;
        const int kind = (t == nt - 1) ? 1 : 0;
        const int cnt = kind ? pl.cnt_last : SDRB_TB;
        double2 v = ypr[k];
        if (pl.correct_iq) {
            const double2 s = make_double2(-offt[t].x, -offt[t].y);
            v = cfma(s, pl.psiY[((size_t)kind * pl.R + r) * SDRB_TB + l], v);
        }
        v = cmul(T1r[t], v);
        const double2 *Win = carry + (size_t)t * 16, *Tn = carry + (size_t)(t + 1) * 16 + 8;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            v = cfma(pl.RW[(size_t)l * 8 + i], Win[i], v);
            v = cfma(pl.RT[(size_t)(cnt - l) * 8 + i], Tn[i], v);
        }
        if (k >= pl.k_bnd) {
#pragma unroll
            for (int i = 0; i < 8; i++) v = cfma(pl.bnd[(size_t)k * 8 + i], s_zeta[i], v);
        }
        yrow[k] = v;
    }
}

// ----------------------------------------------------------------------------------- k_demod
// Stockham radix-2 pass over n points, twiddles from the plan's table (stride fft_n / (2 Ns)).
__device__ __forceinline__ void fft_pass(const double2 *in, double2 *out, int n, int Ns, bool inverse,
                                         const DevPlan &pl)
{
    const int half = n >> 1;
    const int tstride = pl.fft_n / (2 * Ns);
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
        const int k = j & (Ns - 1);
        double2 w = pl.tw[(size_t)k * tstride];
        if (inverse) w.y = -w.y;
        const double2 a = in[j], b = cmul(w, in[j + half]);
        const int j0 = ((j - k) << 1) + k;
        out[j0] = cadd(a, b);
        out[j0 + Ns] = csub(a, b);
    }
}

__device__ __forceinline__ double bswap_double(double v)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    uint32_t lo = (uint32_t)u, hi = (uint32_t)(u >> 32);
    u = ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
    return __longlong_as_double((long long)u);
}

// One CTA per (chunk, row): y[M] complex -> out[row][chunk*M .. +M) doubles.
// Generic entry used both by the chain (y from sc.y) and by the module-level operators.
__global__ void __launch_bounds__(256)
k_demod(const __grid_constant__ DevPlan pl, const double2 *__restrict__ yall, double *__restrict__ out,
        double2 *fftglob, double *zglob, int nchunks, int demod, int apply_sos, int be_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int chunk = blockIdx.x / pl.R, r = blockIdx.x % pl.R;
    if (chunk >= nchunks) return;
    const int M = pl.M, h = M >> 1;
    const double2 *y = yall + ((size_t)chunk * pl.R + r) * M;
    double2 *bufA, *bufB;
    double *z;
    if (pl.demod_in_smem) {
        bufA = reinterpret_cast<double2 *>(smem_raw);
        bufB = bufA + M;
        z = reinterpret_cast<double *>(bufB + M);
    } else {
        bufA = fftglob + (size_t)blockIdx.x * 2 * M;
        bufB = bufA + M;
        z = zglob + (size_t)blockIdx.x * M;
    }
    if (demod == 0) {          // fm: demodulation.py:25-38
        for (int i = threadIdx.x; i < h; i += blockDim.x) {
            const double2 a = y[2 * i], b = y[2 * i + 1];
            // a * conj(b)
            const double re = fma(a.x, b.x, a.y * b.y), im = fma(a.y, b.x, -a.x * b.y);
            bufA[i] = make_double2(atan2(im, re), 0.0);
        }
        __syncthreads();
        if (pl.fft_ok) {
            double2 *src = bufA, *dst = bufB;
            for (int Ns = 1; Ns < h; Ns <<= 1) {
                fft_pass(src, dst, h, Ns, false, pl);
                __syncthreads();
                double2 *tmp = src; src = dst; dst = tmp;
            }
            // spectrum of the 2x interpolated row (scipy.signal.resample: halve the unpaired bin)
            for (int k = threadIdx.x; k < M; k += blockDim.x) {
                double2 v = make_double2(0.0, 0.0);
                const int hh = h >> 1;
                if (k < hh) v = src[k];
                else if (k == hh) v = cscale(0.5, (h > 1) ? src[hh] : src[0]);
                else if (k > M - hh) v = src[h - (M - k)];
                else if (k == M - hh && hh > 0) v = cscale(0.5, cconj(src[hh]));
                dst[k] = v;
            }
            __syncthreads();
            // dst holds Y (length M); inverse FFT ping-pongs between dst and src... src has only
            // M entries when it is bufA/bufB of size M: both buffers are M long.
            double2 *s2 = dst, *d2 = src;
            for (int Ns = 1; Ns < M; Ns <<= 1) {
                fft_pass(s2, d2, M, Ns, true, pl);
                __syncthreads();
                double2 *tmp = s2; s2 = d2; d2 = tmp;
            }
            const double sc1 = 1.0 / (double)h;
            for (int k = threadIdx.x; k < M; k += blockDim.x) z[k] = s2[k].x * sc1;
        } else {
            // dense interpolation matrix built by the host from scipy.signal.resample
            for (int k = threadIdx.x; k < M; k += blockDim.x) {
                const double *row = pl.fm_interp + (size_t)k * h;
                double acc = 0.0;
                for (int i = 0; i < h; i++) acc = fma(row[i], bufA[i].x, acc);
                z[k] = acc;
            }
        }
    } else if (demod == 1) {   // am: abs(square(z)) (demodulation.py:41-48)
        for (int k = threadIdx.x; k < M; k += blockDim.x) {
            const double2 a = y[k];
            z[k] = hypot(fma(a.x, a.x, -a.y * a.y), 2.0 * a.x * a.y);
        }
    } else if (demod == 2) {
        for (int k = threadIdx.x; k < M; k += blockDim.x) z[k] = y[k].x;
    } else {
        for (int k = threadIdx.x; k < M; k += blockDim.x) z[k] = y[k].y;
    }
    __syncthreads();
    // output low-pass: scipy.signal.sosfilt, zero initial state, SciPy's operation order
    if (apply_sos && pl.nsec_out > 0 && threadIdx.x == 0) {
        double z0[4] = {0, 0, 0, 0}, z1[4] = {0, 0, 0, 0};
        for (int k = 0; k < M; k++) {
            double xc = z[k];
            for (int s = 0; s < pl.nsec_out; s++) {
                const double *cf = pl.out_sos + 6 * s;
                const double xn = __dadd_rn(__dmul_rn(cf[0], xc), z0[s]);
                z0[s] = __dadd_rn(__dadd_rn(__dmul_rn(cf[1], xc), -__dmul_rn(cf[4], xn)), z1[s]);
                z1[s] = __dadd_rn(__dmul_rn(cf[2], xc), -__dmul_rn(cf[5], xn));
                xc = xn;
            }
            z[k] = xc;
        }
    }
    __syncthreads();
    double *o = out + ((size_t)r * nchunks + chunk) * M;
    for (int k = threadIdx.x; k < M; k += blockDim.x) o[k] = be_out ? bswap_double(z[k]) : z[k];
}

// demodulation.py:71-79  res[m,n] = y[n] * shift[m,n]
__global__ void k_shift(const double2 *__restrict__ y, const double2 *__restrict__ shift,
                        double2 *__restrict__ res, int R, int N)
{
    const size_t total = (size_t)R * N;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const double2 a = y[idx % N], b = shift[idx];
        res[idx] = make_double2(__dadd_rn(__dmul_rn(a.x, b.x), -__dmul_rn(a.y, b.y)),
                                __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
    }
}
