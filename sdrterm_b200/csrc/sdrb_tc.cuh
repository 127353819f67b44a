// sdrb_tc.cuh -- k_tc: the block front end on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Every per-block quantity of the chain -- decode, byte order, block-local IQ correction, NCO and
// the 16 modal block sums F_i, G_i plus the IQ-EMA block aggregate E -- is a real-linear functional
// of the block's integer samples, i.e. one row of a coefficient matrix applied to the block's raw
// bytes.  sdrterm_b200/plan.py (build_tc) rounds the coefficients to 48-bit fixed point and cuts
// them into balanced base-256 digits, so that
//
//     D[block, NCOL*o + t] = sum_k  rawbyte[block, k] * digit[o, t, k]        (int8 x int8 -> int32)
//
// is an EXACT integer GEMM: A = the raw stream itself, viewed as [blocks][K = q*2*itemsize] bytes
// and brought in by TMA (128B swizzle) without ever being decoded; B = the digit matrix, resident
// in shared memory; D in tensor memory.  The epilogue warps (lane <-> TMEM lane <-> block) read
// their rows with tcgen05.ld, recombine the digit columns in int64 -> FP64 (exact up to one
// rounding) and continue exactly like k_main: tile-local IQ offsets, modal scans in the rotating
// frame, partial outputs and tile aggregates for k_fixup.
//
// Warp roles (640 threads, one CTA per SM, persistent over MMA tiles of 128 blocks = 4 tiles; a
// CTA serves one row r of the VFO bank, blockIdx.x % R):
//   warp 0      TMA producer (one lane)
//   warp 1      TMEM allocation; tcgen05.mma issue (one lane); the MMA's own completion frees the
//               A stage (tcgen05.commit), nothing in the epilogue touches the staged bytes
//   warps 2-3   sign fix-up: XOR 0x80 into the bytes that are not the signed top byte, so that
//               every byte is a valid two's-complement int8 operand (the constant this removes is
//               added back as cst[o])
//   warps 4-19  epilogue: warp = 4 + 8*stage + 4*half + quarter.  `quarter` (= warp%4) is the TMEM
//               lane quarter = one tile of 32 blocks (lane <-> block); half 0 takes the eight
//               forward modal sums F_i (and x0), half 1 the eight backward sums G_i of the same
//               tile; the two meet once per tile on a named barrier to form the partial output
//
// Reference behaviour reproduced: src/misc/read_file.py:100-103, src/dsp/demodulation.py:71-79,
// the block form of scipy.signal.decimate (src/dsp/dsp_processor.py:147).  tests/emulator.py
// (emu_main_tc) is the numpy twin.
#pragma once
#include <cuda.h>
#include "sdrb_kernels.cuh"

#ifndef TC_ABL
#define TC_ABL 0                       // timing experiments only (wrong results): see microbench/ablate.sh
#endif
#define TC_NOUT 36                     // outputs per row: 16 F, 16 G, 2 E, 2 x0
#define TC_MAX_R 32                    // rows of the VFO bank this kernel takes
#define TC_THREADS 640
#define TC_EPI_WARPS 16
#define TC_XS 33                       // exchange-buffer row stride in double2
#define TC_STAGES 2                    // TMEM accumulator stages
#define TC_ASTAGES 3                   // shared-memory A stages (TMA -> sign fix-up -> MMA ring)
#define TC_REGION_BYTES 16384          // 128 rows x 128 bytes, one SWIZZLE_128B operand slab

struct TcDev {
    int K, isz, ncol, nout, npad, nregion;
    uint32_t xor_word;                 // XOR pattern of 4 consecutive stream bytes
    uint32_t idesc;                    // tcgen05 instruction descriptor (i8 x i8 -> s32, M=128, N=npad)
    double scale;                      // 2^-S, common to every fixed-point output
    double scale16;                    // 2^(16-S): weight of the columns above the low pair
    uint32_t stagger_ns;               // one-off delay of accumulator stage 1's first epilogue
    const double2 *prot_pow;           // [R][16][9] powers 0..8 of the rotating-frame block multipliers
    // per row, read through the constant cache (shared-memory bandwidth is this kernel's limit):
    double cstb[TC_MAX_R][TC_NOUT];    // response to the constant the XOR removed, minus the 2^52 + 2^31
                                       // bias of the low digit pair (x0: minus the bias itself)
    double2 phi[TC_MAX_R][16];         // PhiF (0..7), PhiG (8..15)
};

__host__ __device__ inline size_t tc_warp_bytes()
{
    return (size_t)8 * TC_XS * sizeof(double2);                // xs[8 modes][33]
}
__host__ __device__ inline size_t tc_smem_bytes(int npad, int nregion)
{
    size_t b = (size_t)nregion * npad * 128;                   // B slabs
    b += (size_t)TC_ASTAGES * nregion * TC_REGION_BYTES;       // A stages
    b += TC_EPI_WARPS * tc_warp_bytes();                       // xs
    b += 2 * 8 * 32 * sizeof(double2);                         // F/G pair exchange, double-buffered
    b += 16 * 9 * sizeof(double2);                             // prot_pow of this row
    b += 24 * sizeof(unsigned long long);                      // mbarriers
    return b + 1024;                                           // alignment slack
}

// ------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t cnt)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(cnt));
}
__device__ __forceinline__ void mbar_arrive(uint32_t a)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or
// the hint (ns) expires, instead of the warp burning issue slots in a spin loop.
__device__ __forceinline__ bool mbar_try(uint32_t a, uint32_t parity, uint32_t hint_ns)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(a), "r"(parity), "r"(hint_ns) : "memory");
    return ok != 0;
}
// Bounded waits: a protocol error traps (and surfaces as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity)
{
    uint32_t n = 0;
    while (!mbar_try(a, parity, 20000u))
        if (++n > (1u << 22)) __trap();
}
__device__ __forceinline__ void mbar_wait_sleep(uint32_t a, uint32_t parity)
{
    uint32_t n = 0;
    while (!mbar_try(a, parity, 100000u))
        if (++n > (1u << 20)) __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory operand descriptor (8-row atoms of 128 bytes, 1024 B apart).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                  // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset between 8-row atoms
    d |= (uint64_t)1 << 46;                  // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Exact int64 -> double for |v| < 2^51 without the slow I2F.F64 path: bias into the mantissa of
// 1.5 * 2^52 and subtract it again.
__device__ __forceinline__ double i64_to_double(long long v)
{
    return __longlong_as_double(v + 0x4338000000000000LL) - 6755399441055744.0;
}
// int32 -> 2^52 + 2^31 + v as a double (exact); the caller folds the bias into its constant.
__device__ __forceinline__ double i32_biased(int v)
{
    return __hiloint2double(0x43300000, v ^ 0x80000000);
}
#define TC_BIAS32 4503601774854144.0      // 2^52 + 2^31

// Digit columns c[0..NCOL) (most significant first, weight 256 per step) -> value * 2^-S + cst.
// Columns are summed in adjacent pairs in 32 bits (|column| < 2^22), the high pairs are joined by
// one wide multiply-add and converted through the 2^52 bias; the low pair rides on the bias of a
// single FMA whose constant `cstb` = cst - (2^52 + 2^31) * 2^-S already removes it.
template <int NCOL>
__device__ __forceinline__ double tc_combine(const uint32_t *c, double s16, double s1, double cstb)
{
    static_assert(NCOL >= 5 && NCOL <= 7, "digit columns");
    constexpr int NH = NCOL - 2;                       // columns above the low pair
    const int lo = (int)c[NCOL - 2] * 256 + (int)c[NCOL - 1];
    long long hi;
    if (NH == 3) {
        hi = (long long)(int)c[0] * 65536 + (long long)((int)c[1] * 256 + (int)c[2]);
    } else if (NH == 4) {
        hi = (long long)((int)c[0] * 256 + (int)c[1]) * 65536 + (long long)((int)c[2] * 256 + (int)c[3]);
    } else {
        hi = ((long long)((int)c[0] * 256 + (int)c[1]) * 65536 + (long long)((int)c[2] * 256 + (int)c[3])) * 256 +
             (long long)(int)c[4];
    }
    return fma(i64_to_double(hi), s16, fma(i32_biased(lo), s1, cstb));
}

#define TC_DBG(sc, it, ev) do { if ((sc).dbg && blockIdx.x == 0 && (it) < 32) (sc).dbg[(it) * 16 + (ev)] = clock64(); } while (0)

// ------------------------------------------------------------------------------------- k_tc
struct TcShared {
    unsigned char *sB, *sA;
    double2 *sXS, *sPB, *sPow;
    uint32_t bar0;
};

// Epilogue of one warp: HALF 0 = forward modal sums F_i (+ x0, tile aggregate of the IQ offsets),
// HALF 1 = backward sums G_i.  See the file header for the stages.
template <bool IQ, int NCOL, int HALF>
__device__ __forceinline__ void tc_epilogue(const DevPlan &pl, const TcDev &tc, const Scratch &sc, const TcShared &sh,
                                            uint32_t tmem_base, int e, int lane, int r, int slot, int nslots,
                                            int my_iters, int total_tiles)
{
    enum { B_MMA_DONE = 2, B_TMEM_FREE = 4 };
    const int qd = e & 3, g = e >> 3;
    double2 *xs = sh.sXS + (size_t)e * 8 * TC_XS;
    const int pole = lane & 7, seg = lane >> 3;
    const double2 *pw = sh.sPow + (HALF * 8 + pole) * 9;
    const double2 Pm = pw[1], Pm2 = pw[2], Pm8 = pw[8], Pm16 = cmul(pw[8], pw[8]);
    const double2 rot = pl.T3[(size_t)r * (SDRB_TB + 1) + lane];
    const double2 rot31 = pl.T3[(size_t)r * (SDRB_TB + 1) + SDRB_TB - 1];
    const double2 epsb = cconj(pl.T3[(size_t)r * (SDRB_TB + 1) + 1]);
    const uint32_t pair_bar = 1u + (uint32_t)(g * 4 + qd);
    const double s16 = tc.scale16, s1 = tc.scale;
    const uint32_t bar_done = sh.bar0 + 8u * (uint32_t)(B_MMA_DONE * TC_ASTAGES + g);
    const uint32_t bar_free = sh.bar0 + 8u * (uint32_t)(B_TMEM_FREE * TC_ASTAGES + g);
    const uint32_t trow = tmem_base + ((uint32_t)(32 * qd) << 16) + (uint32_t)(g * 256);
    double2 *xsl = xs + lane;                      // lane <-> block view
    const int nt_shift = (pl.ntiles & (pl.ntiles - 1)) == 0 ? 31 - __clz(pl.ntiles) : -1;
    double2 *xsp = xs + pole * TC_XS + seg * 8;    // (segment, mode) view

    for (int it = g; it < my_iters; it += 2) {
        const int u = it >> 1;
        const int mt = slot + it * nslots;
        const int gt = 4 * mt + qd;
        // one warp of the stage polls the mbarrier; the other seven sleep on a named barrier
        // (bar.sync blocks in hardware, a try_wait loop spends issue slots)
        if (HALF == 0 && qd == 0) {
            mbar_wait(bar_done, u & 1);
            // stagger the two accumulator stages by half an item once: left alone they fall into
            // lockstep (both in their integer-heavy TMEM phase, then both in their FP64 scans)
            if (g == 1 && it == 1 && tc.stagger_ns) __nanosleep(tc.stagger_ns);
        }
        asm volatile("bar.sync %0, 256;" ::"r"(9u + (uint32_t)g) : "memory");
        tc_fence_after();
        if (HALF == 0 && qd == 0 && lane == 0) TC_DBG(sc, it, 5);
        if (gt >= total_tiles) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_free);
            continue;
        }
        const int chunk = nt_shift >= 0 ? (gt >> nt_shift) : gt / pl.ntiles, t = gt - chunk * pl.ntiles;

        // ---- TMEM reads are double-buffered: the load of the next 16 columns is in flight while
        //      the current ones are recombined (tcgen05.wait::ld waits for all earlier loads, so
        //      every wait is followed at once by the issue of the next load)
        uint32_t ce[16], ca[16], cb[16];
        constexpr int MODE_COLS = 2 * NCOL;                    // one complex mode = re, im digit columns
        const uint32_t tmode = trow + MODE_COLS * (8 * HALF);
        if (IQ || HALF == 0) tmem_ld16(trow + 32 * NCOL, ce);
        tmem_ld16(tmode, ca);
        tmem_ld_wait();
        tmem_ld16(tmode + MODE_COLS, cb);
        // ---- IQ-EMA block aggregate E -> tile-local block offsets, tile aggregate; x0
        double2 excl = make_double2(0.0, 0.0), x0 = make_double2(0.0, 0.0);
        if (IQ || HALF == 0) {
            if (HALF == 0) {                      // x0: isz exact columns per component
                int xr, xi;
                if (tc.isz == 2) {
                    xr = (int)ce[2 * NCOL] * 256 + (int)ce[2 * NCOL + 1];
                    xi = (int)ce[2 * NCOL + 2] * 256 + (int)ce[2 * NCOL + 3];
                } else {
                    xr = (int)ce[2 * NCOL]; xi = (int)ce[2 * NCOL + 1];
                }
                x0 = make_double2(i32_biased(xr) + tc.cstb[r][34], i32_biased(xi) + tc.cstb[r][35]);
                if (sc.x0) sc.x0[((size_t)chunk * pl.R + r) * pl.Mf + (size_t)t * SDRB_TB + lane] = x0;
            }
            if (IQ) {
                const double er = tc_combine<NCOL>(ce, s16, s1, tc.cstb[r][32]);
                const double ei = tc_combine<NCOL>(ce + NCOL, s16, s1, tc.cstb[r][33]);
                double2 inc = make_double2(pl.Liq * er, pl.Liq * ei);
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    const double2 tt = shfl_up_c(inc, 1 << i);
                    if (lane >= (1 << i)) { inc.x = fma(pl.lamq_pow[i], tt.x, inc.x); inc.y = fma(pl.lamq_pow[i], tt.y, inc.y); }
                }
                excl = shfl_up_c(inc, 1);
                if (lane == 0) excl = make_double2(0.0, 0.0);
                if (HALF == 0 && lane == 31 && r == 0) sc.tile_agg[(size_t)chunk * pl.ntiles + t] = inc;
            }
        }
        if (HALF == 0 && qd == 0 && lane == 0) TC_DBG(sc, it, 8);       // excl known
        // ---- this half's eight modal block sums: digit columns -> FP64, minus the response to
        //      the tile-local offset; lane <-> block, parked mode-major for the scans
#pragma unroll
        for (int md = 0; md < 8; md++) {
            const uint32_t *c = (md & 1) ? cb : ca;
            const int o = 16 * HALF + 2 * md;                  // output index of the real part
            double vr = tc_combine<NCOL>(c, s16, s1, tc.cstb[r][o]);
            double vi = tc_combine<NCOL>(c + NCOL, s16, s1, tc.cstb[r][o + 1]);
            if (md < 7) {
                tmem_ld_wait();                                // mode md + 1 has landed ...
                if (md < 6) tmem_ld16(tmode + MODE_COLS * (md + 2), (md & 1) ? cb : ca);   // ... fetch md + 2
            }
            if (IQ) {
                const double2 ph = tc.phi[r][8 * HALF + md];
                vr = fma(-excl.x, ph.x, fma(excl.y, ph.y, vr));
                vi = fma(-excl.x, ph.y, fma(-excl.y, ph.x, vi));
            }
            if (!(TC_ABL & 16) || md == 0) xsl[md * TC_XS] = make_double2(vr, vi);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free);
        if (HALF == 0 && qd == 0 && lane == 0) TC_DBG(sc, it, 6);

        // ---- tile-local scans in the rotating frame, lane = (segment of 8 blocks, mode):
        //      A. each segment from a zero state, B. carries across the 4 segments,
        //      C. carry applied with the multiplier powers
        // (shared-memory bandwidth is the scarce resource of this kernel -- MMA operand reads, the
        // TMA writes, the sign fix-up and these transposes all share 128 B/clk -- so the scans
        // touch xs exactly once per direction: read for A, write after C)
        double2 loc[8];
        double2 st = make_double2(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double2 v = (TC_ABL & 16) ? make_double2((double)j, excl.x) : xsp[HALF ? 7 - j : j];
            const double2 nst = cfma(Pm, st, v);
            loc[j] = HALF ? nst : st;                              // local (segment-relative) prefix
            st = nst;
        }
        if (HALF == 0 && qd == 0 && lane == 0) TC_DBG(sc, it, 9);       // phase A done
        // carry into this segment from the ones before (after) it: Pm8^2 e_a2 + Pm8 e_a1 + e_a0,
        // two dependent complex FMAs instead of a three-step chain
        double2 cin;
        {
            const int dist = HALF ? 3 - seg : seg;                         // segments feeding this one
            const int dir = HALF ? 1 : -1;
            const double2 n1 = shfl_c(st, ((seg + dir) & 3) * 8 + pole);   // nearest
            const double2 n2 = shfl_c(st, ((seg + 2 * dir) & 3) * 8 + pole);
            const double2 n3 = shfl_c(st, ((seg + 3 * dir) & 3) * 8 + pole);
            const double2 z = make_double2(0.0, 0.0);
            const double2 t1 = cfma(Pm8, dist >= 2 ? n2 : z, dist >= 1 ? n1 : z);
            cin = dist >= 3 ? cfma(Pm16, n3, t1) : t1;
        }
        if (seg == (HALF ? 0 : 3) && !(TC_ABL & 1)) {
            const double2 tot = cfma(Pm8, cin, st);
            double2 *ag = sc.agg + (((size_t)chunk * pl.R + r) * pl.ntiles + t) * 16;
            if (HALF == 0) ag[pole] = cmul(rot31, tot);
            else ag[8 + pole] = tot;
        }
        // ---- C. carry applied with the multiplier powers (table in shared memory, fetched four at a
        //      time so that the loads are in flight together), parked back for the per-block sums
#pragma unroll
        for (int j0 = 0; j0 < 8; j0 += 4) {
            double2 pp[4];
#pragma unroll
            for (int u4 = 0; u4 < 4; u4++) pp[u4] = pw[HALF ? j0 + u4 + 1 : j0 + u4];
#pragma unroll
            for (int u4 = 0; u4 < 4; u4++) {
                const double2 w = cfma(pp[u4], cin, loc[j0 + u4]);
                if (TC_ABL & 8) loc[j0 + u4] = w; else xsp[HALF ? 7 - (j0 + u4) : j0 + u4] = w;
            }
        }
        __syncwarp();
        if (HALF == 0 && qd == 0 && lane == 0) TC_DBG(sc, it, 10);      // phases B, C done
        // ---- partial output of each block (lane <-> block): half 0 sums rho_i W_i, half 1
        //      rho_i/p_i T_i, four independent partial sums
        double2 acc;
        {
            const double2 *rh = HALF ? pl.rho_p : pl.rho;
            const double2 q0 = (TC_ABL & 8) ? cfma(rh[4], loc[4], cmul(rh[0], loc[0])) : cfma(rh[4], xsl[4 * TC_XS], cmul(rh[0], xsl[0]));
            const double2 q1 = (TC_ABL & 8) ? cfma(rh[5], loc[5], cmul(rh[1], loc[1])) : cfma(rh[5], xsl[5 * TC_XS], cmul(rh[1], xsl[1 * TC_XS]));
            const double2 q2 = (TC_ABL & 8) ? cfma(rh[6], loc[6], cmul(rh[2], loc[2])) : cfma(rh[6], xsl[6 * TC_XS], cmul(rh[2], xsl[2 * TC_XS]));
            const double2 q3 = (TC_ABL & 8) ? cfma(rh[7], loc[7], cmul(rh[3], loc[3])) : cfma(rh[7], xsl[7 * TC_XS], cmul(rh[3], xsl[3 * TC_XS]));
            acc = cadd(cadd(q0, q1), cadd(q2, q3));
        }
        double2 *pb = sh.sPB + ((size_t)(u & 1) * 8 + (g * 4 + qd)) * 32;
        if (HALF && !(TC_ABL & 2)) pb[lane] = acc;
        if (HALF == 0 && qd == 0 && lane == 0) TC_DBG(sc, it, 11);      // dot done
        if (!(TC_ABL & 2)) asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        if (HALF == 0 && qd == 0 && lane == 0) TC_DBG(sc, it, 12);      // partner arrived
        if (HALF == 0) {
            const double2 sT = (TC_ABL & 2) ? acc : pb[lane];
            const double2 xc = csub(x0, excl);
            double2 ys = cfma(epsb, acc, sT);
            ys.x = fma(pl.g0, xc.x, ys.x); ys.y = fma(pl.g0, xc.y, ys.y);
            { const double2 yo = cmul(rot, ys);
              if (!(TC_ABL & 4) || yo.x != yo.x) sc.ypart[((size_t)chunk * pl.R + r) * pl.Mf + (size_t)t * SDRB_TB + lane] = yo; }
        }
        __syncwarp();
        if (HALF == 0 && qd == 0 && lane == 0) TC_DBG(sc, it, 7);
    }
}

template <bool IQ, int NCOL>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc(const __grid_constant__ DevPlan pl, const __grid_constant__ TcDev tc, const __grid_constant__ Scratch sc,
     const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
     int nchunks, int n_mtiles)
{
    extern __shared__ unsigned char smem_dyn[];
    // keep the shared address space visible to the compiler: offset the array, do not re-cast
    unsigned char *smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nreg = tc.nregion, npad = tc.npad;
    TcShared sh;
    sh.sB = smem;
    sh.sA = sh.sB + (size_t)nreg * npad * 128;
    sh.sXS = reinterpret_cast<double2 *>(sh.sA + (size_t)TC_ASTAGES * nreg * TC_REGION_BYTES);
    sh.sPB = sh.sXS + (size_t)TC_EPI_WARPS * 8 * TC_XS;
    sh.sPow = sh.sPB + 2 * 8 * 32;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sh.sPow + 16 * 9);
    __shared__ uint32_t tmem_base_s;
    const uint32_t bar0 = smem_u32(bars);
    sh.bar0 = bar0;
    unsigned char *sA = sh.sA, *sB = sh.sB;
    auto BAR = [&](int kind, int s) { return bar0 + 8u * (uint32_t)(kind * TC_ASTAGES + s); };
    enum { B_FULL_A = 0, B_XORED = 1, B_MMA_DONE = 2, B_A_FREE = 3, B_TMEM_FREE = 4, B_BFULL = 5 };

    // one row of the bank per CTA; the CTAs of a row share its MMA tiles round-robin
    const int r = (int)blockIdx.x % pl.R;
    const int slot = (int)blockIdx.x / pl.R, nslots = (int)gridDim.x / pl.R;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_ASTAGES; s++) {
            mbar_init(BAR(B_FULL_A, s), 1);
            mbar_init(BAR(B_XORED, s), 2);
            mbar_init(BAR(B_A_FREE, s), 1);
        }
        for (int s = 0; s < TC_STAGES; s++) {
            mbar_init(BAR(B_MMA_DONE, s), 1);
            mbar_init(BAR(B_TMEM_FREE, s), 8);
        }
        mbar_init(BAR(B_BFULL, 0), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 16 * 9; i += blockDim.x) sh.sPow[i] = tc.prot_pow[(size_t)r * 16 * 9 + i];
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int my_iters = slot < nslots ? (n_mtiles - slot + nslots - 1) / nslots : 0;
    const int total_tiles = nchunks * pl.ntiles;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0 && my_iters > 0) {
            mbar_expect_tx(BAR(B_BFULL, 0), (uint32_t)(nreg * npad * 128));
            for (int rg = 0; rg < nreg; rg++)
                tma_load_2d(smem_u32(sB + (size_t)rg * npad * 128), &map_b, rg * 128, r * npad, BAR(B_BFULL, 0));
            for (int it = 0; it < my_iters; it++) {
                const int s = it % TC_ASTAGES, u = it / TC_ASTAGES;
                const int mt = slot + it * nslots;
                mbar_wait_sleep(BAR(B_A_FREE, s), (u & 1) ^ 1);
                TC_DBG(sc, it, 0);
                mbar_expect_tx(BAR(B_FULL_A, s), (uint32_t)(nreg * TC_REGION_BYTES));
                for (int rg = 0; rg < nreg; rg++)
                    tma_load_2d(smem_u32(sA + ((size_t)s * nreg + rg) * TC_REGION_BYTES), &map_a, rg * 128, mt * 128,
                                BAR(B_FULL_A, s));
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0 && my_iters > 0) {
            mbar_wait_sleep(BAR(B_BFULL, 0), 0);
            for (int it = 0; it < my_iters; it++) {
                const int s = it % TC_ASTAGES, ua = it / TC_ASTAGES;     // A stage
                const int ts = it & 1, u = it >> 1;                      // TMEM stage
                mbar_wait_sleep(BAR(B_XORED, s), ua & 1);
                mbar_wait(BAR(B_TMEM_FREE, ts), (u & 1) ^ 1);
                tc_fence_after();
                TC_DBG(sc, it, 3);
                const uint32_t d = tmem_base + (uint32_t)(ts * 256);
                const int ksteps = tc.K >> 5;
                for (int ks = 0; ks < ksteps; ks++) {
                    const int rg = ks >> 2, kin = (ks & 3) * 32;
                    const uint64_t da = umma_desc(smem_u32(sA + ((size_t)s * nreg + rg) * TC_REGION_BYTES) + kin);
                    const uint64_t db = umma_desc(smem_u32(sB + (size_t)rg * npad * 128) + kin);
                    umma_i8(d, da, db, tc.idesc, ks > 0 ? 1u : 0u);
                }
                umma_commit(BAR(B_A_FREE, s));       // the staged bytes are dead once the MMAs retire
                umma_commit(BAR(B_MMA_DONE, ts));
                TC_DBG(sc, it, 4);
            }
        }
    } else if (warp < 4) {
        // ===================================================================== sign fix-up
        const int tx = threadIdx.x - 64;
        const uint32_t m = tc.xor_word;
        for (int it = 0; it < my_iters; it++) {
            const int s = it % TC_ASTAGES, u = it / TC_ASTAGES;
            mbar_wait_sleep(BAR(B_FULL_A, s), u & 1);
            if (threadIdx.x == 64) TC_DBG(sc, it, 1);
            uint4 *base = reinterpret_cast<uint4 *>(sA + (size_t)s * nreg * TC_REGION_BYTES);
            const int n16 = nreg * (TC_REGION_BYTES / 16);
#pragma unroll 8
            for (int i = tx; i < n16; i += 64) {
                uint4 v = base[i];
                v.x ^= m; v.y ^= m; v.z ^= m; v.w ^= m;
                base[i] = v;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_XORED, s));
            if (threadIdx.x == 64) TC_DBG(sc, it, 2);
        }
    } else {
        // ===================================================================== epilogue
        const int e = warp - 4;
        if ((e >> 2) & 1) tc_epilogue<IQ, NCOL, 1>(pl, tc, sc, sh, tmem_base, e, lane, r, slot, nslots, my_iters, total_tiles);
        else tc_epilogue<IQ, NCOL, 0>(pl, tc, sc, sh, tmem_base, e, lane, r, slot, nslots, my_iters, total_tiles);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}
