// sdrb_kernels.cuh -- the sm_100a kernels of the demodulation chain.
//
//   k_main     raw tile -> shared memory; run-wise IQ correction, NCO and the even/odd block sums
//              on the FP64 tensor pipe (DMMA.8x8x4), all from registers; tile-local modal scans
//              -> partial outputs + tile aggregates
//   k_iqgain / k_iqscan / k_iqtiles   IQ-corrector offset at every tile start, carried across
//              chunks and calls
//   k_fixup    head / end segments, cross-tile carries, boundary term -> decimated complex y
//   k_demod    fm (pair phase + 2x FFT interpolation) | am | re | im, segmented output SOS, framing
//
// Reference behaviour being reproduced: src/misc/read_file.py:100-103 (decode, normalise, IQ
// correction), src/dsp/demodulation.py:71-79 (NCO), src/dsp/dsp_processor.py:147 (scipy
// decimate), demodulation.py:25-68, dsp_processor.py:32-36,149,162, vfo_processor.py:84.
// tests/emulator.py is the numpy twin of these kernels (same tables, same decomposition).
#pragma once
#include "sdrb_device.cuh"

// Scratch written by k_main / read by k_fixup, per batch of chunks.
struct Scratch {
    double2 *ypart;     // [nch][R][Mf]
    double2 *agg;       // [nch][R][ntiles][16]   0..7 Wout, 8..15 Tin
    double2 *tile_agg;  // [nch][ntiles]
    double2 *off_tile;  // [nch][ntiles+1]
    double2 *gain;      // [nch] offset gained over a chunk from zero (whole-tile fast path)
    double2 *start;     // [nch] offset at the chunk start            "
    double2 *carry;     // [nch][R][ntiles+1][16] 0..7 Win[t], 8..15 Tn[t]
    double2 *y;         // [nch][R][M]
    double2 *iq_state;  // [1] offset before the batch (in) / after it (out)
    double2 *fftbuf;    // [nch*R][2][M] when the demod buffers do not fit shared memory
    double *zrow;       // [nch*R][M]     "
    unsigned long long *dbg;  // optional k_tc timeline (SDRB_TC_DEBUG): [64 tiles][16 events] clock64 of CTA 0
    double2 *x0;        // optional [nch][R][Mf]: first raw sample of every block as the tensor-core
                        // front end sees it (exact integers; sdrb_keep_x0, parity tests)
};

// Shared-memory carve-up of k_main (host mirrors this in sdrb_api.cu).
// A raw tile is kept in the order the DMMA fragments consume it: slot (s, g, lane) holds the two
// raw samples lane = (block 8g + (lane>>2), pair kq*RL + s) multiplies in step s of group g --
// sample j and its mirror q-1-j -- side by side, so a lane fetches a pair with ONE shared-memory
// load of 2*sb bytes, conflict-free, and no address arithmetic on the row.  Planes of one s are
// 256*sb bytes apart plus 16 bytes of skew (the staging stores of a warp walk along s).
__host__ __device__ inline size_t main_plane_bytes(int sb) { return (size_t)256 * sb + 16; }
__host__ __device__ inline size_t main_tile_bytes(int RL, int sb)
{
    return (size_t)RL * main_plane_bytes(sb) + 2 * SDRB_TB * sizeof(double2);   // pairs + cl[32], blkagg[32]
}
__host__ __device__ inline size_t main_warp_bytes()
{
    return (size_t)32 * SDRB_XSTRIDE * sizeof(double) + SDRB_TB * sizeof(double2);  // xb + x0s
}

// Raw tiles of one CTA -> shared memory in fragment order, one word of type T per sample: every
// warp takes four block rows at a time and issues their loads together (rows past the chunk's
// last block are zero).
// Stored byte order -> the GPU's, one raw sample (two items) at a time: done once when the tile is
// staged, not once per row of the bank.
__device__ __forceinline__ uint16_t swap_items(uint16_t v) { return v; }
__device__ __forceinline__ uint32_t swap_items(uint32_t v) { return __byte_perm(v, 0, 0x2301); }
__device__ __forceinline__ uint2 swap_items(uint2 v) { return make_uint2(__byte_perm(v.x, 0, 0x0123), __byte_perm(v.y, 0, 0x0123)); }
__device__ __forceinline__ uint4 swap_items(uint4 v)
{
    return make_uint4(__byte_perm(v.y, 0, 0x0123), __byte_perm(v.x, 0, 0x0123), __byte_perm(v.w, 0, 0x0123), __byte_perm(v.z, 0, 0x0123));
}

template <typename T>
__device__ __forceinline__ void stage_pairs(unsigned char *tiles, size_t tileb, const uint8_t *rawc, int t0, int ntl,
                                            int ntiles, int cnt_last, int q, int Hq, int RL, int swap, int warp, int W, int lane)
{
    const int total = ntl * SDRB_TB;
    const size_t plane = main_plane_bytes((int)sizeof(T));
    for (int r0 = warp * 4; r0 < total; r0 += W * 4) {
        const int tl = r0 >> 5, row0 = r0 & 31, t = t0 + tl;          // four rows never straddle a tile
        const int cnt = (t == ntiles - 1) ? cnt_last : SDRB_TB;
        const T *src = reinterpret_cast<const T *>(rawc) + ((size_t)t * SDRB_TB + row0) * q;
        unsigned char *tb = tiles + (size_t)tl * tileb;
        for (int col = lane; col < q; col += 32) {
            const int slot = col < Hq ? 0 : 1, jj = slot ? q - 1 - col : col;
            const int kq = jj / RL, sp = jj - kq * RL;
            T v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = (row0 + u < cnt) ? src[(size_t)u * q + col] : T{};
            if (swap) {
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = swap_items(v[u]);
            }
            T *dst = reinterpret_cast<T *>(tb + (size_t)sp * plane) + ((size_t)(row0 >> 3) * 32 + (row0 & 7) * 4 + kq) * 2 + slot;
#pragma unroll
            for (int u = 0; u < 4; u++) dst[u * 8] = v[u];           // next block row: lane slot + 4 = 8 words on
        }
    }
    // slots no sample lands in -- pairs Hq .. 4 RL - 1 and the mirror of an odd block's middle
    // sample -- are zero: their modal coefficients are zero, so the data only has to be finite
    const int nempty = 2 * (4 * RL - Hq) + (q & 1);
    for (int idx = warp * 32 + lane; idx < nempty * SDRB_TB * ntl; idx += W * 32) {
        const int b = idx & 31, k = (idx >> 5) % nempty, tl = (idx >> 5) / nempty;
        const int jj = k < 2 * (4 * RL - Hq) ? Hq + (k >> 1) : Hq - 1, slot = k < 2 * (4 * RL - Hq) ? (k & 1) : 1;
        const int kq = jj / RL, sp = jj - kq * RL;
        reinterpret_cast<T *>(tiles + (size_t)tl * tileb + (size_t)sp * plane)[((size_t)(b >> 3) * 32 + (b & 7) * 4 + kq) * 2 + slot] = T{};
    }
}

// The pair of slot (s, g, lane) -> two complex doubles (decode of read_file.py:100-101 from
// registers; the byte order was fixed when the tile was staged).
template <int ENC>
__device__ __forceinline__ void load_pair(const DevPlan &pl, const unsigned char *plane_s, int g, int lane, double2 &za, double2 &zb)
{
    const int idx = g * 32 + lane;
    if (ENC == ENC_b || ENC == ENC_B) {
        const uint32_t v = reinterpret_cast<const uint32_t *>(plane_s)[idx];
        if (ENC == ENC_b) {
            za = make_double2((double)(int8_t)(v & 0xff), (double)(int8_t)((v >> 8) & 0xff));
            zb = make_double2((double)(int8_t)((v >> 16) & 0xff), (double)(int8_t)(v >> 24));
        } else {
            za = make_double2((double)(v & 0xff), (double)((v >> 8) & 0xff));
            zb = make_double2((double)((v >> 16) & 0xff), (double)(v >> 24));
        }
    } else if (ENC == ENC_h || ENC == ENC_H) {
        const uint2 v = reinterpret_cast<const uint2 *>(plane_s)[idx];
        if (ENC == ENC_h) {
            za = make_double2((double)(int16_t)(v.x & 0xffff), (double)(int16_t)(v.x >> 16));
            zb = make_double2((double)(int16_t)(v.y & 0xffff), (double)(int16_t)(v.y >> 16));
        } else {
            za = make_double2((double)(v.x & 0xffff), (double)(v.x >> 16));
            zb = make_double2((double)(v.y & 0xffff), (double)(v.y >> 16));
        }
    } else if (ENC == ENC_i || ENC == ENC_I || ENC == ENC_f) {
        const uint4 v = reinterpret_cast<const uint4 *>(plane_s)[idx];
        if (ENC == ENC_i) {
            za = make_double2((double)(int32_t)v.x, (double)(int32_t)v.y); zb = make_double2((double)(int32_t)v.z, (double)(int32_t)v.w);
        } else if (ENC == ENC_I) {
            za = make_double2((double)v.x, (double)v.y); zb = make_double2((double)v.z, (double)v.w);
        } else {
            za = make_double2((double)__uint_as_float(v.x), (double)__uint_as_float(v.y));
            zb = make_double2((double)__uint_as_float(v.z), (double)__uint_as_float(v.w));
        }
    } else {
        const unsigned char *p = plane_s + (size_t)idx * 32;
        za = load_sample<ENC>(p, 0, 0);
        zb = load_sample<ENC>(p, 1, 0);
    }
    if (pl.normalize) {
        za = normalize_sample(za, pl.norm_xmin, pl.norm_k);
        zb = normalize_sample(zb, pl.norm_xmin, pl.norm_k);
    }
}

// ------------------------------------------------------------------------------------ k_main
// grid.x = nchunks * ceil(ntiles / TPC); block = 32*W threads.  A CTA stages TPC consecutive
// tiles of one chunk in shared memory (raw bytes, once, shared by all rows); its warps then take
// (tile, row) items.  Lane roles follow the DMMA fragments: k = lane&3 owns the pairs
// [k*RL, (k+1)*RL) of block 8g + (lane>>2) in group g (B operand), and accumulates output row
// lane>>2 for blocks 8g + 2k, 8g + 2k + 1 (C operand).
template <int ENC, bool IQ>
__global__ void __launch_bounds__(256, 2)
k_main(const __grid_constant__ DevPlan pl, Scratch sc, const uint8_t *__restrict__ raw, int nchunks, int TPC)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int q = pl.q, sb = pl.sb, RL = pl.RL;
    const int groups = (pl.ntiles + TPC - 1) / TPC;
    const int chunk = blockIdx.x / groups, tg = blockIdx.x % groups;
    if (chunk >= nchunks) return;
    const size_t tileb = main_tile_bytes(RL, sb), plane = main_plane_bytes(sb);
    unsigned char *warp_base = smem_raw + (size_t)TPC * tileb;
    const uint8_t *rawc = raw + (size_t)chunk * pl.N * sb;
    const int t0 = tg * TPC;
    const int ntl = min(TPC, pl.ntiles - t0);
    const int kq = lane & 3, nq = lane >> 2;
    // this lane's pair range inside a block: pairs [a0, a1), i.e. samples j and q-1-j
    const int a0 = min(kq * RL, pl.Hq), a1 = min((kq + 1) * RL, pl.Hq);
    const int npair = a1 - a0;                              // valid steps s < npair
    const int smid = (q & 1) && a1 == pl.Hq ? npair - 1 : -1;   // the step that holds the middle sample of an odd block

    // ---------------- stage the raw tiles (all warps, four rows in flight per lane), then (IQ) run
    //                  aggregates and block offsets
    if (sb == 2) stage_pairs<uint16_t>(smem_raw, tileb, rawc, t0, ntl, pl.ntiles, pl.cnt_last, q, pl.Hq, RL, pl.swap, warp, W, lane);
    else if (sb == 4) stage_pairs<uint32_t>(smem_raw, tileb, rawc, t0, ntl, pl.ntiles, pl.cnt_last, q, pl.Hq, RL, pl.swap, warp, W, lane);
    else if (sb == 8) stage_pairs<uint2>(smem_raw, tileb, rawc, t0, ntl, pl.ntiles, pl.cnt_last, q, pl.Hq, RL, pl.swap, warp, W, lane);
    else stage_pairs<uint4>(smem_raw, tileb, rawc, t0, ntl, pl.ntiles, pl.cnt_last, q, pl.Hq, RL, pl.swap, warp, W, lane);
    __syncthreads();
    for (int tl = warp; tl < ntl; tl += W) {
        const int t = t0 + tl;
        const int cnt = (t == pl.ntiles - 1) ? pl.cnt_last : SDRB_TB;
        unsigned char *tb = smem_raw + (size_t)tl * tileb;
        double2 *cl = reinterpret_cast<double2 *>(tb + (size_t)RL * plane);
        double2 *blkagg = cl + SDRB_TB;
        double2 excl = make_double2(0.0, 0.0);
        if (IQ) {
            // IQ-EMA aggregate of every block (the samples themselves stay uncorrected: the
            // corrector enters through the decoupled terms)
            for (int g = 0; g < 4; g++) {
                const int b = 8 * g + nq;
                double2 agg_a = make_double2(0.0, 0.0), agg_d = make_double2(0.0, 0.0);
                // ascending run = the first samples of the pairs (Horner); descending run = their
                // mirrors q-1-j, newest last: sum_sp lam^sp zb_sp (one decode pass for both)
                double wd = 1.0;
                for (int sp = 0; sp < npair; sp++) {
                    double2 za, zb;
                    load_pair<ENC>(pl, tb + (size_t)sp * plane, g, lane, za, zb);
                    agg_a.x = fma(pl.lam, agg_a.x, za.x); agg_a.y = fma(pl.lam, agg_a.y, za.y);
                    if (sp != smid) { agg_d.x = fma(wd, zb.x, agg_d.x); agg_d.y = fma(wd, zb.y, agg_d.y); }
                    wd *= pl.lam;
                }
                // EMA state at the 9 run boundaries of this block (all four lanes of the block
                // compute the same chain from the gathered run aggregates)
                const int base = lane & ~3;
                double2 A[9];
                A[0] = make_double2(0.0, 0.0);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const double2 v = shfl_c(agg_a, base + i);
                    A[i + 1] = make_double2(fma(pl.lam_run[i], A[i].x, v.x), fma(pl.lam_run[i], A[i].y, v.y));
                }
#pragma unroll
                for (int i = 4; i < 8; i++) {
                    const double2 v = shfl_c(agg_d, base + (7 - i));
                    A[i + 1] = make_double2(fma(pl.lam_run[i], A[i].x, v.x), fma(pl.lam_run[i], A[i].y, v.y));
                }
                if (kq == 0) blkagg[b] = A[8];
            }
            __syncwarp();
            const bool active = lane < cnt;
            double2 inc = active ? cscale(pl.Liq, blkagg[lane]) : make_double2(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < 5; i++) {
                const double2 tt = shfl_up_c(inc, 1 << i);
                if (lane >= (1 << i)) { inc.x = fma(pl.lamq_pow[i], tt.x, inc.x); inc.y = fma(pl.lamq_pow[i], tt.y, inc.y); }
            }
            excl = shfl_up_c(inc, 1);
            if (lane == 0) excl = make_double2(0.0, 0.0);
            const double2 tagg = shfl_c(inc, cnt - 1);
            if (lane == 0) sc.tile_agg[(size_t)chunk * pl.ntiles + t] = tagg;
        }
        cl[lane] = excl;
    }
    __syncthreads();

    // ---------------- items: (tile, row)
    double *xb = reinterpret_cast<double *>(warp_base + (size_t)warp * main_warp_bytes());
    double2 *x0s = reinterpret_cast<double2 *>(xb + 32 * SDRB_XSTRIDE);
    const int e = (lane >> 2) & 1, m = lane >> 3;
    for (int item = warp; item < ntl * pl.R; item += W) {
        const int tl = item % ntl, r = item / ntl;
        const int t = t0 + tl;
        const int cnt = (t == pl.ntiles - 1) ? pl.cnt_last : SDRB_TB;
        const unsigned char *tb = smem_raw + (size_t)tl * tileb;
        const double2 *cl = reinterpret_cast<const double2 *>(tb + (size_t)RL * plane);
        const double2 *T2r = pl.T2 + (size_t)r * q;
        const bool nco = pl.use_nco[r] != 0;

        // step s outside, the four block groups inside: the NCO phasors of the pair and the modal
        // A fragments are fetched once per step and serve 4 x 4 DMMAs
        double acc[4][8];
#pragma unroll
        for (int g = 0; g < 4; g++)
#pragma unroll
            for (int i = 0; i < 8; i++) acc[g][i] = 0.0;
        for (int sp = 0; sp < RL; sp++) {
            // steps past this lane's pairs (and the odd part of a middle sample) carry zero modal
            // coefficients and zero data: no selects in the loop
            const int j = a0 + sp;
            double2 Ta = make_double2(1.0, 0.0), Tb = Ta;
            if (nco && sp < npair) { Ta = __ldg(T2r + j); Tb = __ldg(T2r + (q - 1 - j)); }
            const double aE = __ldg(pl.Afrag + (size_t)(2 * sp) * 32 + lane);
            const double aO = __ldg(pl.Afrag + (size_t)(2 * sp + 1) * 32 + lane);
            const unsigned char *ps = tb + (size_t)sp * plane;
            const bool killb = pl.normalize && sp == smid;         // (a normalised zero is not zero)
            const bool first = sp == 0 && kq == 0;
#pragma unroll
            for (int g = 0; g < 4; g++) {
                double2 za, zb;
                load_pair<ENC>(pl, ps, g, lane, za, zb);
                if (first) x0s[8 * g + nq] = za;                   // first (raw) sample of the block
                if (killb) zb = make_double2(0.0, 0.0);
                const double2 ua = nco ? cmul(Ta, za) : za, ub = nco ? cmul(Tb, zb) : zb;
                const double2 a = cadd(ua, ub), d = csub(ua, ub);
                dmma884(acc[g][0], acc[g][1], aE, a.x);
                dmma884(acc[g][2], acc[g][3], aE, a.y);
                dmma884(acc[g][4], acc[g][5], aO, d.x);
                dmma884(acc[g][6], acc[g][7], aO, d.y);
            }
        }
        // combine the four real sums into F/G (part e of poles m and m+4), exchange
#pragma unroll
        for (int g = 0; g < 4; g++) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int b = 8 * g + 2 * kq + i;
                const double Sr = acc[g][i], Si = acc[g][2 + i], Dr = acc[g][4 + i], Di = acc[g][6 + i];
                const double oS = __shfl_xor_sync(0xffffffffu, Si, 4);
                const double oD = __shfl_xor_sync(0xffffffffu, Di, 4);
                double Sup, Slo, Dup, Dlo;
                if (e == 0) { Sup = Sr - oS; Slo = Sr + oS; Dup = Dr - oD; Dlo = Dr + oD; }
                else        { Sup = oS + Sr; Slo = oS - Sr; Dup = oD + Dr; Dlo = oD - Dr; }
                const double Fu = Sup + Dup, Gu = Sup - Dup, Fl = Slo + Dlo, Gl = Slo - Dlo;
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
            // forward lanes: exclusive scan = the inclusive one stored one slot up (slot 0 = 0, slot
            // cnt is padding or an unused block); backward lanes: inclusive, in place.  The next
            // input is fetched before the store that may overwrite it.
            const int dir = fwd ? 1 : -1;
            double *xr = xb + (2 * lane) * SDRB_XSTRIDE + (fwd ? 0 : cnt - 1), *xi = xr + SDRB_XSTRIDE;
            double *orr = xr + (fwd ? 1 : 0), *oi = orr + SDRB_XSTRIDE;
            double2 v = make_double2(*xr, *xi);
            if (fwd) { *xr = 0.0; *xi = 0.0; }
            for (int step = 0; step < cnt; step++) {
                xr += dir; xi += dir;
                double2 vn = make_double2(0.0, 0.0);
                if (step + 1 < cnt) vn = make_double2(*xr, *xi);
                st = cfma(Pm, st, v);
                *orr = st.x; *oi = st.y;
                orr += dir; oi += dir;
                v = vn;
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
                sw = cfma(pl.rb[(size_t)r * 8 + i], Wv, sw);
                sT = cfma(pl.rbT[(size_t)r * 8 + i], Tv, sT);
            }
            const double2 epsb = cconj(pl.T3[(size_t)r * (SDRB_TB + 1) + 1]);
            const double2 x0 = x0s[lane];
            double2 ys = cfma(epsb, sw, sT);
            ys.x = fma(pl.g0, x0.x, ys.x); ys.y = fma(pl.g0, x0.y, ys.y);
            if (IQ) {                                  // - gamma * (tile-local offset at this block)
                const double2 gm = pl.gamma[r], o = cl[lane];
                ys = cfma(make_double2(-gm.x, -gm.y), o, ys);
            }
            const double2 yp = cmul(pl.T3[(size_t)r * (SDRB_TB + 1) + lane], ys);
            sc.ypart[((size_t)chunk * pl.R + r) * pl.Mf + (size_t)t * SDRB_TB + lane] = yp;
        }
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------- IQ kernels
// k_iqgain: thread <-> chunk, offset gained over the chunk from a zero state -> off_tile[c][nt].
template <int ENC>
__global__ void __launch_bounds__(128)
k_iqgain(const __grid_constant__ DevPlan pl, Scratch sc, const uint8_t *__restrict__ raw, int nchunks)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const int nt = pl.ntiles;
    const double2 *ta = sc.tile_agg + (size_t)c * nt;
    double2 o = make_double2(0.0, 0.0);
    int t = 0;
    for (; t + 8 <= nt; t += 8) {
        double2 g[8];
#pragma unroll
        for (int i = 0; i < 8; i++) g[i] = ta[t + i];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const double lt = pl.lam_tile[(t + i) == nt - 1 ? 1 : 0];
            o.x = fma(lt, o.x, g[i].x); o.y = fma(lt, o.y, g[i].y);
        }
    }
    for (; t < nt; t++) {
        const double lt = pl.lam_tile[t == nt - 1 ? 1 : 0];
        const double2 g = ta[t];
        o.x = fma(lt, o.x, g.x); o.y = fma(lt, o.y, g.y);
    }
    if (pl.rem) {
        const uint8_t *rawc = raw + (size_t)c * pl.N * pl.sb;
        double2 acc = make_double2(0.0, 0.0);
        for (int j = 0; j < pl.rem; j++) {
            const double2 z = decode_sample<ENC>(pl, rawc, (long)pl.q * pl.Mf + j);
            acc.x = fma(pl.lam, acc.x, z.x); acc.y = fma(pl.lam, acc.y, z.y);
        }
        const double lr = pl.lam_j[pl.rem];
        o.x = fma(lr, o.x, pl.Liq * acc.x); o.y = fma(lr, o.y, pl.Liq * acc.y);
    }
    sc.off_tile[(size_t)c * (nt + 1) + nt] = o;
}

// k_iqgain_w: warp <-> chunk, lane <-> tile (whole-tile chunks, ntiles <= 32): the same gain as
// k_iqgain by a warp scan of the tile maps o -> lam_tile o + agg[t]; coalesced, no serial loop.
__global__ void __launch_bounds__(256)
k_iqgain_w(const __grid_constant__ DevPlan pl, Scratch sc, int nchunks)
{
    pdl_trigger();
    pdl_wait();                                   // tile_agg comes from the block front end
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= nchunks) return;
    const int nt = pl.ntiles;
    double m = 1.0;
    double2 a = make_double2(0.0, 0.0);
    if (lane < nt) { m = pl.lam_tile[lane == nt - 1 ? 1 : 0]; a = sc.tile_agg[(size_t)c * nt + lane]; }
#pragma unroll
    for (int lv = 0; lv < 5; lv++) {
        const double pm = __shfl_up_sync(0xffffffffu, m, 1 << lv);
        const double2 pa = shfl_up_c(a, 1 << lv);
        if (lane >= (1 << lv)) { a.x = fma(m, pa.x, a.x); a.y = fma(m, pa.y, a.y); m *= pm; }
    }
    if (lane == nt - 1) sc.gain[c] = a;
}

// k_iqchunk: warp <-> chunk: the offset a whole chunk gains from a zero state, straight from the raw
// bytes (decode, normalise, EMA) -> gain[c].  The pre-pass of time-segment sharding for streams
// longer than one batch (sdrb_iq_gain): no outputs, the raw bytes are read once.
template <int ENC>
__global__ void __launch_bounds__(256)
k_iqchunk(const __grid_constant__ DevPlan pl, Scratch sc, const uint8_t *__restrict__ raw, int nchunks)
{
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= nchunks) return;
    const uint8_t *rawc = raw + (size_t)c * pl.N * pl.sb;
    // lane <-> contiguous run of ceil(N/32) samples
    const int per = (pl.N + 31) / 32;
    const int n0 = min(pl.N, lane * per), n1 = min(pl.N, n0 + per);
    double2 a = make_double2(0.0, 0.0);
    double m = 1.0;
    for (int n = n0; n < n1; n++) {
        const double2 z = decode_sample<ENC>(pl, rawc, n);
        a.x = fma(pl.lam, a.x, z.x); a.y = fma(pl.lam, a.y, z.y);
        m *= pl.lam;
    }
#pragma unroll
    for (int lv = 0; lv < 5; lv++) {
        const double pm = __shfl_up_sync(0xffffffffu, m, 1 << lv);
        const double2 pa = shfl_up_c(a, 1 << lv);
        if (lane >= (1 << lv)) { a.x = fma(m, pa.x, a.x); a.y = fma(m, pa.y, a.y); m *= pm; }
    }
    if (lane == 31) sc.gain[c] = make_double2(pl.Liq * a.x, pl.Liq * a.y);
}

// k_iqscan_c: one CTA of 1024 threads; exclusive scan of the per-chunk maps o -> lam_N o + gain[c]
// from the handle's IQ state -> start[c] and the new state.  8 chunks per thread and round (their
// gains are loaded together), warp-shuffle scans, one shared-memory hop across the 32 warps.
__global__ void __launch_bounds__(1024)
k_iqscan_c(const __grid_constant__ DevPlan pl, Scratch sc, int nchunks)
{
    __shared__ double2 s_a[32];
    __shared__ double s_m[32];
    __shared__ double2 s_state;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double2 *__restrict__ gain = sc.gain;
    double2 *__restrict__ start = sc.start;
    pdl_trigger();
    pdl_wait();                                   // gain[] comes from k_iqgain_w / k_iqchunk
    if (tid == 0) s_state = sc.iq_state[0];
    __syncthreads();
    for (int base = 0; base < nchunks; base += 1024 * 8) {
        const int c0 = base + tid * 8;
        double2 g[8];
#pragma unroll
        for (int i = 0; i < 8; i++) g[i] = (c0 + i < nchunks) ? gain[c0 + i] : make_double2(0.0, 0.0);
        double m = 1.0;
        double2 a = make_double2(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (c0 + i < nchunks) { a.x = fma(pl.lam_N, a.x, g[i].x); a.y = fma(pl.lam_N, a.y, g[i].y); m *= pl.lam_N; }
        // inclusive scan of the thread maps inside the warp
#pragma unroll
        for (int lv = 0; lv < 5; lv++) {
            const double pm = __shfl_up_sync(0xffffffffu, m, 1 << lv);
            const double2 pa = shfl_up_c(a, 1 << lv);
            if (lane >= (1 << lv)) { a.x = fma(m, pa.x, a.x); a.y = fma(m, pa.y, a.y); m *= pm; }
        }
        if (lane == 31) { s_a[warp] = a; s_m[warp] = m; }
        __syncthreads();
        if (warp == 0) {
            double wm = s_m[lane];
            double2 wa = s_a[lane];
#pragma unroll
            for (int lv = 0; lv < 5; lv++) {
                const double pm = __shfl_up_sync(0xffffffffu, wm, 1 << lv);
                const double2 pa = shfl_up_c(wa, 1 << lv);
                if (lane >= (1 << lv)) { wa.x = fma(wm, pa.x, wa.x); wa.y = fma(wm, pa.y, wa.y); wm *= pm; }
            }
            s_a[lane] = wa; s_m[lane] = wm;
        }
        __syncthreads();
        // exclusive map of this thread = (warps before) then (lanes before in this warp)
        double em = __shfl_up_sync(0xffffffffu, m, 1);
        double2 ea = shfl_up_c(a, 1);
        if (lane == 0) { em = 1.0; ea = make_double2(0.0, 0.0); }
        if (warp > 0) {
            const double wm = s_m[warp - 1];
            const double2 wa = s_a[warp - 1];
            ea.x = fma(em, wa.x, ea.x); ea.y = fma(em, wa.y, ea.y); em *= wm;
        }
        const double2 st0 = s_state;
        double2 o = make_double2(fma(em, st0.x, ea.x), fma(em, st0.y, ea.y));
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (c0 + i < nchunks) {
                start[c0 + i] = o;
                o.x = fma(pl.lam_N, o.x, g[i].x); o.y = fma(pl.lam_N, o.y, g[i].y);
            }
        __syncthreads();
        if (tid == 1023) s_state = make_double2(fma(s_m[31], st0.x, s_a[31].x), fma(s_m[31], st0.y, s_a[31].y));
        __syncthreads();
    }
    if (tid == 0) sc.iq_state[0] = s_state;
}

// k_iqscan: one CTA; exclusive scan of the per-chunk affine maps o -> lam_N*o + gain, starting from
// the handle's IQ state; writes the offset at every chunk start to off_tile[c][0] and the new state.
__global__ void __launch_bounds__(1024)
k_iqscan(const __grid_constant__ DevPlan pl, Scratch sc, int nchunks)
{
    __shared__ double2 s_a[1024];
    __shared__ double s_m[1024];
    const int tid = threadIdx.x, nth = blockDim.x, nt = pl.ntiles;
    const int per = (nchunks + nth - 1) / nth;
    const int c0 = min(nchunks, tid * per), c1 = min(nchunks, c0 + per);
    double2 a = make_double2(0.0, 0.0);
    double mlt = 1.0;
    for (int c = c0; c < c1; c++) {
        const double2 g = sc.off_tile[(size_t)c * (nt + 1) + nt];
        a.x = fma(pl.lam_N, a.x, g.x); a.y = fma(pl.lam_N, a.y, g.y);
        mlt *= pl.lam_N;
    }
    s_a[tid] = a; s_m[tid] = mlt;
    __syncthreads();
    // Hillis-Steele inclusive scan of affine maps (m, a): later o earlier = (m2*m1, m2*a1 + a2)
    for (int d = 1; d < nth; d <<= 1) {
        double2 pa = make_double2(0.0, 0.0);
        double pm = 1.0;
        const bool has = tid >= d;
        if (has) { pa = s_a[tid - d]; pm = s_m[tid - d]; }
        __syncthreads();
        if (has) {
            const double mm = s_m[tid];
            s_a[tid] = make_double2(fma(mm, pa.x, s_a[tid].x), fma(mm, pa.y, s_a[tid].y));
            s_m[tid] = mm * pm;
        }
        __syncthreads();
    }
    const double2 st0 = sc.iq_state[0];
    double2 o = st0;
    if (tid > 0) {
        const double2 pa = s_a[tid - 1]; const double pm = s_m[tid - 1];
        o = make_double2(fma(pm, st0.x, pa.x), fma(pm, st0.y, pa.y));
    }
    for (int c = c0; c < c1; c++) {
        sc.off_tile[(size_t)c * (nt + 1)] = o;
        const double2 g = sc.off_tile[(size_t)c * (nt + 1) + nt];
        o.x = fma(pl.lam_N, o.x, g.x); o.y = fma(pl.lam_N, o.y, g.y);
    }
    __syncthreads();
    if (tid == nth - 1) {
        const double2 pa = s_a[tid]; const double pm = s_m[tid];
        sc.iq_state[0] = make_double2(fma(pm, st0.x, pa.x), fma(pm, st0.y, pa.y));
    }
}

// k_iqtiles: thread <-> chunk, offsets at the tile starts from the chunk-start offset.
__global__ void __launch_bounds__(128)
k_iqtiles(const __grid_constant__ DevPlan pl, Scratch sc, int nchunks)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const int nt = pl.ntiles;
    const double2 *ta = sc.tile_agg + (size_t)c * nt;
    double2 *ot = sc.off_tile + (size_t)c * (nt + 1);
    double2 o = ot[0];
    int t = 0;
    for (; t + 8 <= nt; t += 8) {
        double2 g[8];
#pragma unroll
        for (int i = 0; i < 8; i++) g[i] = ta[t + i];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            ot[t + i] = o;
            const double lt = pl.lam_tile[(t + i) == nt - 1 ? 1 : 0];
            o.x = fma(lt, o.x, g[i].x); o.y = fma(lt, o.y, g[i].y);
        }
    }
    for (; t < nt; t++) {
        ot[t] = o;
        const double lt = pl.lam_tile[t == nt - 1 ? 1 : 0];
        const double2 g = ta[t];
        o.x = fma(lt, o.x, g.x); o.y = fma(lt, o.y, g.y);
    }
    ot[nt] = o;      // offset at sample q*Mf
}

// Time-segment sharding: the IQ offset a segment gained from a zero start, and the
// fold of the gains of the segments before this rank's into its start offset -- on the device, so
// that the exchange between the two passes needs no host round trip.
__global__ void k_iq_export(Scratch sc, double *dst3, double nsamples)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double2 g = sc.iq_state[0];
        dst3[0] = g.x; dst3[1] = g.y; dst3[2] = nsamples;
    }
}
__global__ void k_iq_prefix(Scratch sc, const double *gains3, int rank, double lam)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double2 off = make_double2(0.0, 0.0);
        for (int i = 0; i < rank; i++) {
            const double d = pow(lam, gains3[3 * i + 2]);
            off.x = fma(d, off.x, gains3[3 * i]); off.y = fma(d, off.y, gains3[3 * i + 1]);
        }
        sc.iq_state[0] = off;
    }
}

// ----------------------------------------------------------------------------------- k_fixup
// One CTA per (chunk, row).  Warp 0: raw head / end-window samples are fetched by all lanes at
// once, then lanes 0..7 <-> poles run the head, the cross-tile carries, the end segment and the
// boundary vector zeta; then all threads emit y[k].
template <int ENC>
__global__ void __launch_bounds__(128)
k_fixup(const __grid_constant__ DevPlan pl, Scratch sc, const uint8_t *__restrict__ raw, int nchunks)
{
    __shared__ double2 s_x[320];       // head samples (edge+1) then end-window samples (nend)
    __shared__ double2 s_zeta[SDRB_NP];
    const int chunk = blockIdx.x / pl.R, r = blockIdx.x % pl.R;
    if (chunk >= nchunks) return;
    const int nt = pl.ntiles, edge = pl.edge, q = pl.q;
    const uint8_t *rawc = raw + (size_t)chunk * pl.N * pl.sb;
    const double2 *offt = sc.off_tile + (size_t)chunk * (nt + 1);
    const double2 *aggr = sc.agg + ((size_t)chunk * pl.R + r) * nt * 16;
    double2 *carry = sc.carry + ((size_t)chunk * pl.R + r) * (nt + 1) * 16;
    double2 *yrow = sc.y + ((size_t)chunk * pl.R + r) * pl.M;
    const double2 *T1r = pl.T1 + (size_t)r * nt;
    double2 *s_h = s_x, *s_e = s_x + (edge + 1);
    const bool iq = pl.correct_iq != 0;

    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        // fetch: raw head samples and the raw end window; every load is independent
        for (int n = lane; n <= edge; n += 32) s_h[n] = decode_sample<ENC>(pl, rawc, n);
        for (int i = lane; i < pl.nend; i += 32) s_e[i] = decode_sample<ENC>(pl, rawc, (long)pl.ws + i);
        __syncwarp();
        // serial IQ recurrences over these few samples (read_file.py:72-77): the head runs forward
        // from the chunk-start offset, the partial block forward from the offset at q*Mf, and the
        // window samples inside full blocks BACKWARD from the offset at q*Mf
        // (off[n] = (off[n+1] - L z[n]) / lam), then the NCO phases
        if (iq) {
            if (lane == 0) {
                double2 o = offt[0];
                for (int n = 0; n <= edge; n++) {
                    const double2 x = csub(s_h[n], o);
                    o.x = fma(x.x, pl.Liq, o.x); o.y = fma(x.y, pl.Liq, o.y);
                    s_h[n] = x;
                }
            } else if (lane == 1 && pl.rem) {
                double2 o = offt[nt];
                for (int i = pl.nend - pl.rem; i < pl.nend; i++) {
                    const double2 x = csub(s_e[i], o);
                    o.x = fma(x.x, pl.Liq, o.x); o.y = fma(x.y, pl.Liq, o.y);
                    s_e[i] = x;
                }
            } else if (lane == 2) {
                double2 o = offt[nt];
                for (int i = pl.nend - pl.rem - 1; i >= 0; i--) {
                    const double2 z = s_e[i];
                    o.x = fma(-pl.Liq, z.x, o.x) * pl.lam_inv; o.y = fma(-pl.Liq, z.y, o.y) * pl.lam_inv;
                    s_e[i] = csub(z, o);
                }
            }
            __syncwarp();
        }
        for (int n = lane; n <= edge; n += 32) s_h[n] = cmul(s_h[n], pl.Ehead[(size_t)r * (edge + 1) + n]);
        for (int i = lane; i < pl.nend; i += 32) s_e[i] = cmul(s_e[i], pl.Eend[(size_t)r * pl.nend + i]);
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
        // forward carries across tiles, in the frame of the UNcorrected samples the front ends
        // summed: W~ = (w + alpha s) / beta with s = off e^{jwn} (n = 0: the chunk-start offset);
        // the carries are stored pre-scaled by beta for the output stage
        const double2 al = pl.alpha[(size_t)r * 8 + i], be = pl.beta[(size_t)r * 8 + i];
        const double2 alT = pl.alphaT[(size_t)r * 8 + i], beT = pl.betaT[(size_t)r * 8 + i];
        const double2 sE = iq ? cmul(offt[nt], pl.phE[r]) : make_double2(0.0, 0.0);   // rotated offset at q*Mf
        if (iq) w = cmul(cfma(al, offt[0], w), pl.binv[(size_t)r * 8 + i]);
        for (int t = 0; t < nt; t++) {
            if (lane < 8) carry[(size_t)t * 16 + i] = iq ? cmul(be, w) : w;
            const int kind = (t == nt - 1) ? 1 : 0;
            const int cnt = kind ? pl.cnt_last : SDRB_TB;
            w = cfma(pl.Ppow[(size_t)cnt * 8 + i], w, cmul(T1r[t], aggr[(size_t)t * 16 + i]));
        }
        if (iq) w = csub(cmul(be, w), cmul(al, sE));              // back to the true state at q*Mf
        if (lane < 8) carry[(size_t)nt * 16 + i] = iq ? cfma(al, sE, w) : w;   // = beta W~
        const double2 wEnd = w;
        // end segment: partial block then tail extension
        const int nseq = pl.rem + edge;
        const double2 xN1 = s_e[pl.nend - 1];
        double2 wL1 = w, last = make_double2(0.0, 0.0);
        for (int k = 0; k < nseq; k++) {
            const double2 v = (k < pl.rem) ? s_e[pl.nend - pl.rem + k]
                                           : csub(cscale(2.0, xN1), s_e[pl.nend - 2 - (k - pl.rem)]);
            if (k == nseq - 1) { wL1 = w; last = v; }
            w = cfma(p, w, v);
        }
        const double2 wL = w;
        double2 T = make_double2(0.0, 0.0);
        for (int k = nseq - 1; k >= 0; k--) {
            const double2 v = (k < pl.rem) ? s_e[pl.nend - pl.rem + k]
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
            zeta = csub(zeta, cmul(pl.xi[i * 8 + l], wl));
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
        // backward carries: T~ = (T + alphaT s) / betaT at q*Mf, stored pre-scaled by betaT
        T = Tend;
        if (iq) T = cmul(cfma(alT, sE, T), pl.binvT[(size_t)r * 8 + i]);
        if (lane < 8) carry[(size_t)nt * 16 + 8 + i] = iq ? cmul(beT, T) : T;
        for (int t = nt - 1; t >= 0; t--) {
            const int kind = (t == nt - 1) ? 1 : 0;
            const int cnt = kind ? pl.cnt_last : SDRB_TB;
            T = cfma(pl.Ppow[(size_t)cnt * 8 + i], T, cmul(T1r[t], aggr[(size_t)t * 16 + 8 + i]));
            if (lane < 8) carry[(size_t)t * 16 + 8 + i] = iq ? cmul(beT, T) : T;
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
        if (iq) {
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
    // output low-pass (scipy.signal.sosfilt, zero initial state) in 32 segments: each lane of
    // warp 0 runs SciPy's DF2T recurrence over its segment from a zero state, the segment-end
    // states are chained with A^Lseg, and every sample then gets c A^i s_in of its segment.
    if (apply_sos && pl.nsec_out > 0 && threadIdx.x < 32) {
        const int lane = threadIdx.x, Ls = pl.sos_Lseg, ns = pl.sos_ns, nsec = pl.nsec_out;
        const int lo = min(M, lane * Ls), hi = min(M, lo + Ls);
        double st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = lo; k < hi; k++) {
            double xc = z[k];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                if (s < nsec) {
                    const double *cf = pl.out_sos + 6 * s;
                    const double xn = __dadd_rn(__dmul_rn(cf[0], xc), st[2 * s]);
                    st[2 * s] = __dadd_rn(__dadd_rn(__dmul_rn(cf[1], xc), -__dmul_rn(cf[4], xn)), st[2 * s + 1]);
                    st[2 * s + 1] = __dadd_rn(__dmul_rn(cf[2], xc), -__dmul_rn(cf[5], xn));
                    xc = xn;
                }
            }
            z[k] = xc;
        }
        // chain: s_in(seg+1) = A^Ls s_in(seg) + s_loc(seg); every lane walks the chain and keeps
        // the value that belongs to its own segment
        double sin_mine[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cur[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int seg = 0; seg < 32; seg++) {
            if (seg == lane) {
#pragma unroll
                for (int a = 0; a < 8; a++) sin_mine[a] = cur[a];
            }
            double nxt[8];
#pragma unroll
            for (int a = 0; a < 8; a++) {
                double v = (a < ns) ? __shfl_sync(0xffffffffu, st[a], seg) : 0.0;
                if (a < ns)
                    for (int b = 0; b < ns; b++) v = fma(pl.sos_AL[a * ns + b], cur[b], v);
                nxt[a] = v;
            }
#pragma unroll
            for (int a = 0; a < 8; a++) cur[a] = nxt[a];
        }
        for (int k = lo; k < hi; k++) {
            const double *ca = pl.sos_CA + (size_t)(k - lo) * ns;
            double v = z[k];
            for (int a = 0; a < ns; a++) v = fma(ca[a], sin_mine[a], v);
            z[k] = v;
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

// read_file.py:100-101  z = y['re'] + 1j*y['im'] on the structured view of the raw bytes.
template <int ENC>
__global__ void k_decode(const uint8_t *__restrict__ raw, double2 *__restrict__ z, size_t n, int swap)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        z[i] = load_sample<ENC>(raw, (long)i, swap);
}

// read_file.py:65-77 / extra/src/iq_correction.pyx:46-52: z[i] -= off; off += z[i] * L, as the
// linear recurrence off' = lam off + L z evaluated by one CTA: thread <-> contiguous run, affine
// maps scanned across the threads, second pass applies the offsets.
__global__ void __launch_bounds__(1024)
k_correct_iq(double2 *__restrict__ z, size_t n, double2 *__restrict__ off_io, double L)
{
    __shared__ double2 s_a[32];
    __shared__ double s_m[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double lam = 1.0 - L;
    const size_t per = (n + blockDim.x - 1) / blockDim.x;
    const size_t i0 = min(n, (size_t)tid * per), i1 = min(n, i0 + per);
    double2 a = make_double2(0.0, 0.0);
    double m = 1.0;
    for (size_t i = i0; i < i1; i++) {
        const double2 v = z[i];
        a.x = fma(lam, a.x, L * v.x); a.y = fma(lam, a.y, L * v.y);
        m *= lam;
    }
#pragma unroll
    for (int lv = 0; lv < 5; lv++) {
        const double pm = __shfl_up_sync(0xffffffffu, m, 1 << lv);
        const double2 pa = shfl_up_c(a, 1 << lv);
        if (lane >= (1 << lv)) { a.x = fma(m, pa.x, a.x); a.y = fma(m, pa.y, a.y); m *= pm; }
    }
    if (lane == 31) { s_a[warp] = a; s_m[warp] = m; }
    __syncthreads();
    if (warp == 0) {
        double wm = s_m[lane];
        double2 wa = s_a[lane];
#pragma unroll
        for (int lv = 0; lv < 5; lv++) {
            const double pm = __shfl_up_sync(0xffffffffu, wm, 1 << lv);
            const double2 pa = shfl_up_c(wa, 1 << lv);
            if (lane >= (1 << lv)) { wa.x = fma(wm, pa.x, wa.x); wa.y = fma(wm, pa.y, wa.y); wm *= pm; }
        }
        s_a[lane] = wa; s_m[lane] = wm;
    }
    __syncthreads();
    double em = __shfl_up_sync(0xffffffffu, m, 1);
    double2 ea = shfl_up_c(a, 1);
    if (lane == 0) { em = 1.0; ea = make_double2(0.0, 0.0); }
    if (warp > 0) {
        const double wm = s_m[warp - 1];
        const double2 wa = s_a[warp - 1];
        ea.x = fma(em, wa.x, ea.x); ea.y = fma(em, wa.y, ea.y); em *= wm;
    }
    const double2 st0 = off_io[0];
    double2 o = make_double2(fma(em, st0.x, ea.x), fma(em, st0.y, ea.y));
    for (size_t i = i0; i < i1; i++) {
        double2 v = z[i];
        v.x -= o.x; v.y -= o.y;
        o.x = fma(v.x, L, o.x); o.y = fma(v.y, L, o.y);
        z[i] = v;
    }
    __syncthreads();
    if (tid == blockDim.x - 1) off_io[0] = o;
}

// dsp_processor.py:159-160  z[:] = savgol_filter(z, window, 3) per chunk row of M outputs, as the
// three linear maps SciPy uses (tab = head rows | FIR row | tail rows, each `w` long).
__global__ void k_savgol(const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ tab,
                         int w, int nhead, int ntail, int lo, int M, size_t nseg)
{
    for (size_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        const double *xs = x + seg * M;
        double *ys = y + seg * M;
        for (int k = threadIdx.x; k < M; k += blockDim.x) {
            int row, base;
            if (k < nhead) { row = k; base = 0; }
            else if (k >= M - ntail) { row = nhead + 1 + (k - (M - ntail)); base = M - w; }
            else { row = nhead; base = k + lo; }
            const double *sr = tab + (size_t)row * w;
            double acc = 0.0;
            for (int j = 0; j < w; j++) {
                const int i = base + j;
                if (i >= 0 && i < M) acc = fma(sr[j], xs[i], acc);
            }
            ys[k] = acc;
        }
    }
}
