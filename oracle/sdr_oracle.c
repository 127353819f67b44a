/*
 * sdr_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never on the product path).
 *
 * Plain-C restatement of the arithmetic on peads/sdrterm's streamed IQ
 * demodulation path, written from the reference's behaviour (not its text):
 *
 *   decode            src/misc/read_file.py:51,100-101,124   (struct view -> re + 1j*im)
 *   normalize         src/misc/read_file.py:82-96,177-196
 *   IQ correction     src/misc/read_file.py:65-77, extra/src/iq_correction.pyx:46-52
 *   NCO shift         src/dsp/demodulation.py:71-79 (x[m,n] = z[n]*shift[m,n])
 *   decimate          src/dsp/dsp_processor.py:147 -> SciPy 1.18.1 signal.decimate
 *                     (cheby1 SOS, sosfiltfilt: odd extension, sosfilt_zi, DF2T
 *                     recurrence; _signaltools.py:5091-5203, 5206-5369)
 *   fm pair phase     src/dsp/demodulation.py:25-32
 *   am                src/dsp/demodulation.py:41-48
 *   output sosfilt    src/dsp/dsp_processor.py:32-36,149 (zero initial state)
 *
 * SciPy is a third-party dependency of the reference (pyproject.toml:12-18, unpinned;
 * 1.18.1 in this image); its compiled _sosfilt loop is restated here with the operation
 * order that is bit-identical to it (tests/test_oracle_pins.py checks this against
 * scipy.signal.decimate / sosfilt on random data).  Compile with -ffp-contract=off:
 * the shipped SciPy build contains no FMA in that loop.
 *
 * All functions take plain pointers; complex arrays are interleaved (re, im) doubles.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- decode: one raw sample pair -> (re, im) doubles.  enc in "bBhHiIfd", swap=1 when the
 * stored byte order differs from the host's (read_file.py:48-51). */
static inline uint16_t bs16(uint16_t v) { return (uint16_t)((v >> 8) | (v << 8)); }
static inline uint32_t bs32(uint32_t v) { return __builtin_bswap32(v); }
static inline uint64_t bs64(uint64_t v) { return __builtin_bswap64(v); }

int orc_itemsize(char enc)
{
    switch (enc) {
    case 'b': case 'B': return 1;
    case 'h': case 'H': return 2;
    case 'i': case 'I': case 'f': return 4;
    case 'd': return 8;
    default: return -1;
    }
}

static inline double decode_one(const uint8_t *p, char enc, int swap)
{
    switch (enc) {
    case 'b': return (double)(int8_t)p[0];
    case 'B': return (double)p[0];
    case 'h': { uint16_t v; memcpy(&v, p, 2); if (swap) v = bs16(v); return (double)(int16_t)v; }
    case 'H': { uint16_t v; memcpy(&v, p, 2); if (swap) v = bs16(v); return (double)v; }
    case 'i': { uint32_t v; memcpy(&v, p, 4); if (swap) v = bs32(v); return (double)(int32_t)v; }
    case 'I': { uint32_t v; memcpy(&v, p, 4); if (swap) v = bs32(v); return (double)v; }
    case 'f': { uint32_t v; float f; memcpy(&v, p, 4); if (swap) v = bs32(v); memcpy(&f, &v, 4); return (double)f; }
    case 'd': { uint64_t v; double d; memcpy(&v, p, 8); if (swap) v = bs64(v); memcpy(&d, &v, 8); return d; }
    default: return NAN;
    }
}

/* raw bytes -> n complex samples (read_file.py:100-101) */
int orc_decode(const uint8_t *raw, long n, char enc, int swap, double *z)
{
    int sz = orc_itemsize(enc);
    if (sz < 0) return -1;
    for (long i = 0; i < n; i++) {
        z[2 * i] = decode_one(raw + (2 * i) * sz, enc, swap);
        z[2 * i + 1] = decode_one(raw + (2 * i + 1) * sz, enc, swap);
    }
    return 0;
}

/* read_file.py:88-96: res = 1.6*(z - xmin)*xMaxMinDiff - 0.8 on COMPLEX z with real xmin and real
 * constants: the real part gets both offsets, the imaginary part only the scale.
 * Operation order as numba evaluates it: ((1.6*(z-xmin))*k) - 0.8 */
void orc_normalize(double *z, long n, double xmin, double xMaxMinDiff)
{
    for (long i = 0; i < n; i++) {
        double re = z[2 * i] - xmin, im = z[2 * i + 1];
        re = 1.6 * re; im = 1.6 * im;
        re = re * xMaxMinDiff; im = im * xMaxMinDiff;
        z[2 * i] = re - 0.8; z[2 * i + 1] = im;
    }
}

/* read_file.py:72-77 / iq_correction.pyx:46-52:  z[i] -= off; off += z[i]*L  (state carried) */
void orc_correct_iq(double *z, long n, double L, double *off)
{
    double ore = off[0], oim = off[1];
    for (long i = 0; i < n; i++) {
        double re = z[2 * i] - ore, im = z[2 * i + 1] - oim;
        z[2 * i] = re; z[2 * i + 1] = im;
        ore += re * L; oim += im * L;
    }
    off[0] = ore; off[1] = oim;
}

/* demodulation.py:71-79: res[n] = y[n]*shift[n], plain complex multiply */
void orc_shift(const double *z, const double *shift, long n, double *out)
{
    for (long i = 0; i < n; i++) {
        double a = z[2 * i], b = z[2 * i + 1], c = shift[2 * i], d = shift[2 * i + 1];
        out[2 * i] = a * c - b * d;
        out[2 * i + 1] = a * d + b * c;
    }
}

/* SciPy _sosfilt inner loop on a complex signal with real coefficients (DF2T):
 *   xn = b0*xc + z0;  z0 = (b1*xc - a1*xn) + z1;  z1 = b2*xc - a2*xn;  xc = xn
 * sos: nsec x 6 (b0 b1 b2 a0 a1 a2), a0 == 1.  zi: nsec x 2 complex (interleaved), updated. */
void orc_sosfilt_c(const double *sos, int nsec, double *x, long n, double *zi)
{
    for (long i = 0; i < n; i++) {
        double xr = x[2 * i], xi = x[2 * i + 1];
        for (int s = 0; s < nsec; s++) {
            const double *c = sos + 6 * s;
            double *st = zi + 4 * s; /* z0r z0i z1r z1i */
            double yr = c[0] * xr + st[0], yi = c[0] * xi + st[1];
            st[0] = (c[1] * xr - c[4] * yr) + st[2];
            st[1] = (c[1] * xi - c[4] * yi) + st[3];
            st[2] = c[2] * xr - c[5] * yr;
            st[3] = c[2] * xi - c[5] * yi;
            xr = yr; xi = yi;
        }
        x[2 * i] = xr; x[2 * i + 1] = xi;
    }
}

/* same recurrence, real signal (dsp_processor.py:32-36 -> scipy.signal.sosfilt, zero state) */
void orc_sosfilt_r(const double *sos, int nsec, double *x, long n, double *zi)
{
    for (long i = 0; i < n; i++) {
        double xc = x[i];
        for (int s = 0; s < nsec; s++) {
            const double *c = sos + 6 * s;
            double *st = zi + 2 * s;
            double xn = c[0] * xc + st[0];
            st[0] = (c[1] * xc - c[4] * xn) + st[1];
            st[1] = c[2] * xc - c[5] * xn;
            xc = xn;
        }
        x[i] = xc;
    }
}

/* scipy.signal.decimate(x, q) for ftype='iir', zero_phase=True on one complex row of length n:
 * sosfiltfilt (odd extension by `edge`, zi from sosfilt_zi scaled by the first/last sample) and
 * y[::q].  zi0: nsec x 2 real (sosfilt_zi).  work: 2*(n+2*edge) doubles.  out: ceil(n/q) complex. */
long orc_decimate(const double *sos, int nsec, const double *zi0, int edge,
                  const double *x, long n, int q, double *work, double *out)
{
    long L = n + 2 * (long)edge;
    double *ext = work;
    /* odd_ext: left 2*x[0] - x[edge..1], right 2*x[n-1] - x[n-2..n-1-edge] */
    for (int j = 0; j < edge; j++) {
        ext[2 * j] = 2 * x[0] - x[2 * (edge - j)];
        ext[2 * j + 1] = 2 * x[1] - x[2 * (edge - j) + 1];
    }
    memcpy(ext + 2 * edge, x, sizeof(double) * 2 * n);
    for (int j = 0; j < edge; j++) {
        ext[2 * (edge + n + j)] = 2 * x[2 * (n - 1)] - x[2 * (n - 2 - j)];
        ext[2 * (edge + n + j) + 1] = 2 * x[2 * (n - 1) + 1] - x[2 * (n - 2 - j) + 1];
    }
    double zi[4 * 16];
    if (nsec > 16) return -1;
    for (int s = 0; s < nsec; s++) {
        zi[4 * s + 0] = zi0[2 * s] * ext[0]; zi[4 * s + 1] = zi0[2 * s] * ext[1];
        zi[4 * s + 2] = zi0[2 * s + 1] * ext[0]; zi[4 * s + 3] = zi0[2 * s + 1] * ext[1];
    }
    orc_sosfilt_c(sos, nsec, ext, L, zi);
    /* reverse in place */
    for (long i = 0, j = L - 1; i < j; i++, j--) {
        double tr = ext[2 * i], ti = ext[2 * i + 1];
        ext[2 * i] = ext[2 * j]; ext[2 * i + 1] = ext[2 * j + 1];
        ext[2 * j] = tr; ext[2 * j + 1] = ti;
    }
    for (int s = 0; s < nsec; s++) {
        zi[4 * s + 0] = zi0[2 * s] * ext[0]; zi[4 * s + 1] = zi0[2 * s] * ext[1];
        zi[4 * s + 2] = zi0[2 * s + 1] * ext[0]; zi[4 * s + 3] = zi0[2 * s + 1] * ext[1];
    }
    orc_sosfilt_c(sos, nsec, ext, L, zi);
    long m = 0;
    for (long i = 0; i < n; i += q, m++) {
        long r = L - 1 - (edge + i); /* un-reverse */
        out[2 * m] = ext[2 * r]; out[2 * m + 1] = ext[2 * r + 1];
    }
    return m;
}

/* demodulation.py:25-32: res[i/2] = angle(y[i]*conj(y[i+1])) over non-overlapping pairs */
void orc_fm_pairs(const double *y, long m, double *res)
{
    for (long i = 0; i + 1 < m; i += 2) {
        double a = y[2 * i], b = y[2 * i + 1], c = y[2 * i + 2], d = -y[2 * i + 3];
        res[i >> 1] = atan2(a * d + b * c, a * c - b * d);
    }
}

/* demodulation.py:41-48: abs(square(z)) == hypot(re(z^2), im(z^2)) */
void orc_am(const double *y, long m, double *res)
{
    for (long i = 0; i < m; i++) {
        double a = y[2 * i], b = y[2 * i + 1];
        res[i] = hypot(a * a - b * b, a * b + b * a);
    }
}

/* One chunk, R rows: shift (optional) + decimate.  z: n complex (already decoded/corrected);
 * shift: R x n complex table or NULL (no shift, R must be 1); y: R x M complex, M = ceil(n/q).
 * Rows are spread over OpenMP threads when available (the reference runs them serially). */
long orc_chunk_rows(const double *sos, int nsec, const double *zi0, int edge, const double *z,
                    const double *shift, int R, long n, int q, double *y, int nthreads)
{
    long M = (n + q - 1) / q;
    int fail = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic)
#endif
    for (int r = 0; r < R; r++) {
        double *x = (double *)malloc(sizeof(double) * 2 * n);
        double *work = (double *)malloc(sizeof(double) * 2 * (n + 2 * (long)edge));
        if (!x || !work) { fail = 1; free(x); free(work); continue; }
        if (shift) orc_shift(z, shift + 2 * (long)r * n, n, x);
        else memcpy(x, z, sizeof(double) * 2 * n);
        orc_decimate(sos, nsec, zi0, edge, x, n, q, work, y + 2 * (long)r * M);
        free(x); free(work);
    }
    (void)nthreads;
    return fail ? -1 : M;
}

/* Many chunks x R rows in one call, chunk-parallel (the chain is chunk-local once the IQ
 * corrector has run; dsp_processor.py:164-183).  z: nchunks x n complex. y: nchunks x R x M. */
long orc_batch(const double *sos, int nsec, const double *zi0, int edge, const double *z,
               const double *shift, int R, long n, int q, long nchunks, double *y, int nthreads)
{
    long M = (n + q - 1) / q;
    long total = nchunks * R;
    int fail = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
#endif
    {
        double *x = (double *)malloc(sizeof(double) * 2 * n);
        double *work = (double *)malloc(sizeof(double) * 2 * (n + 2 * (long)edge));
        if (!x || !work) fail = 1;
#ifdef _OPENMP
#pragma omp for schedule(dynamic)
#endif
        for (long t = 0; t < total; t++) {
            if (fail) continue;
            long c = t / R; int r = (int)(t % R);
            const double *zc = z + 2 * c * n;
            if (shift) orc_shift(zc, shift + 2 * (long)r * n, n, x);
            else memcpy(x, zc, sizeof(double) * 2 * n);
            orc_decimate(sos, nsec, zi0, edge, x, n, q, work, y + 2 * (c * R + r) * M);
        }
        free(x); free(work);
    }
    (void)nthreads;
    return fail ? -1 : M;
}

/* Batched helpers for the CPU baseline: same arithmetic, chunk/row-parallel where the reference's
 * data flow allows it (decode is element-wise; the IQ recurrence stays serial). */
int orc_decode_batch(const uint8_t *raw, long n_total, char enc, int swap, double *z, int nthreads)
{
    int sz = orc_itemsize(enc);
    if (sz < 0) return -1;
    const long blk = 1 << 16;
    long nblk = (n_total + blk - 1) / blk;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
#endif
    for (long b = 0; b < nblk; b++) {
        long lo = b * blk, hi = lo + blk < n_total ? lo + blk : n_total;
        orc_decode(raw + 2 * lo * sz, hi - lo, enc, swap, z + 2 * lo);
    }
    (void)nthreads;
    return 0;
}

void orc_fm_pairs_rows(const double *y, long rows, long m, double *res, int nthreads)
{
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
#endif
    for (long r = 0; r < rows; r++) orc_fm_pairs(y + 2 * r * m, (m >> 1) * 2, res + r * (m >> 1));
    (void)nthreads;
}

void orc_sosfilt_rows(const double *sos, int nsec, double *x, long rows, long n, int nthreads)
{
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
#endif
    for (long r = 0; r < rows; r++) {
        double zi[2 * 16];
        memset(zi, 0, sizeof zi);
        orc_sosfilt_r(sos, nsec, x + r * n, n, zi);
    }
    (void)nthreads;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
