// age_sweep.cuh -- the wavefront sweep: the main sweep kernels (sweep_window_steady / sweep_window_edge) and the trace
// recomputation of pass 2 (sweep_window_compact).
//
// Reference semantics (src/AsmvarAGE/src/AGEaligner.cpp): K1/K2 calcScores :978-1083 and K3/K4 calcMaxima :862-976.
// One warp sweeps one strip (32 lanes x 8 rows) of one pair along the reference window; lane p works on block
// t-p (4 columns) in step t and hands the H/G/M values of its bottom row to lane p+1 with warp shuffles.
#pragma once
#include "age_device.cuh"
#include <type_traits>

namespace age {

#ifndef AGE_V_PREFETCH
#define AGE_V_PREFETCH 1          // build switch (A/B): compact re-sweep with its loads ahead of their use (pre-sweep of the reverse maxima trace 4.1 -> 3.7 ms per C2 batch)
#endif
enum { MODE_TRACE_S = 1, MODE_TRACE_M = 2 };   // trace re-sweeps: S = FS / RS (scores trace), M = FM / RM (maxima trace) + hot bits

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(short)(v & 0xFFFFu); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)(short)(v >> 16); }

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Lane arithmetic of the recurrence.  Ar<false>: two aligners per word, 2*score + tag in signed 16-bit halves (the fast
// path).  Ar<true> ("wide"): one aligner per word in 32 bits, with the reference's storage type made explicit -- every
// score is stored as `unsigned short` (AGEaligner.hh:86, AGEaligner.cpp:1019,1073), so the doubled value is reduced
// modulo 2^17 after the trace decision (`store`), and the split sum `unsigned short val = *maxf + *maxr` (:479,548)
// modulo 2^17 as well (`sum`); the equality tests of findBPs (:493,565) compare plain ints and are not wrapped.
template <bool WIDE> struct Ar;
template <> struct Ar<false> {
    static constexpr uint32_t TAG = 0x00010001u, NTAG = 0xFFFEFFFEu;
    static __device__ __forceinline__ uint32_t maxs(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
    static __device__ __forceinline__ uint32_t addmax_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2_relu(a, b, c); }
    static __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) { return __vadd2(a, b); }
    static __device__ __forceinline__ uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
    static __device__ __forceinline__ uint32_t maxu(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
    static __device__ __forceinline__ uint32_t store(uint32_t x, uint32_t, uint32_t) { return x; }
    static __device__ __forceinline__ uint32_t sum(uint32_t x) { return x; }
    static __device__ __forceinline__ uint32_t pack(int a, int b) { return (uint32_t)(a & 0xFFFF) | ((uint32_t)b << 16); }
    static __device__ __forceinline__ int lo(uint32_t v) { return (int)(short)(v & 0xFFFFu); }
    static __device__ __forceinline__ int hi(uint32_t v) { return (int)(short)(v >> 16); }
};
template <> struct Ar<true> {
    static constexpr uint32_t TAG = 1u, NTAG = 0xFFFFFFFEu;
    static __device__ __forceinline__ uint32_t maxs(uint32_t a, uint32_t b) { return (uint32_t)max((int)a, (int)b); }
    static __device__ __forceinline__ uint32_t addmax_relu(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)__viaddmax_s32_relu((int)a, (int)b, (int)c); }
    static __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) { return a + b; }
    static __device__ __forceinline__ uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)__vimax3_s32((int)a, (int)b, (int)c); }
    static __device__ __forceinline__ uint32_t maxu(uint32_t a, uint32_t b) { return max(a, b); }
    // what the matrix holds: the u16 store (rowmask = 0x1FFFF on rows of the contig) -- and 0 in padding rows / columns
    // (rowmask or colmask 0), one LOP3.  With gap costs >= 0 padding cells would otherwise grow scores of their own and
    // leak into the reverse sweep, which meets them first; zeroed they behave exactly like the matrix border (H = 0, trace
    // 0, G = gap open).  The packed path needs none of this: its gap costs are negative, so padding never exceeds 0.
    static __device__ __forceinline__ uint32_t store(uint32_t x, uint32_t rowmask, uint32_t colmask) { return x & rowmask & colmask; }
    static __device__ __forceinline__ uint32_t sum(uint32_t x) { return x & 0x1FFFEu; }
    static __device__ __forceinline__ uint32_t pack(int a, int) { return (uint32_t)a; }
    static __device__ __forceinline__ int lo(uint32_t v) { return (int)v; }
    static __device__ __forceinline__ int hi(uint32_t v) { return (int)v; }
};

// ---- bulk asynchronous copies (TMA engine, 1-D form) on an mbarrier -------------------------------------------------
// AGE_DSTAGE_BULK = 1: the reverse sweep stages its D slots with cp.async.bulk (one elected lane, mbarrier); 0: per-lane
// cp.async (LDGSTS).  Measured on a B200 with the bench batch (round 2, profiles/r2_dstage_bulk_vs_ldgsts.md): sweep kernel
// 40.8 ms with LDGSTS, 42.9 ms with bulk copies one step ahead, 44.6 ms two steps ahead (three slots) -- the bulk form
// executes 4 % more instructions (the barrier wait spins where cp.async.wait_group just stalls) and a bulk copy takes
// longer to land than 32 lanes' own 16-byte copies.  Hence the default.
#ifndef AGE_DSTAGE_BULK
#define AGE_DSTAGE_BULK 0
#endif
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
// `bytes` (a multiple of 16) from global to shared memory; completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

// substitution score (x2) of one half for a special column (Scorer.hh:24-39): codes >= 5 score 0
__device__ __forceinline__ int subst2(int xc, int yc, int m2, int mm2)
{
    if (xc >= 5 || yc >= 5) return 0;
    return xc == yc ? m2 : mm2;
}
// OFF8 format of D (age_device.cuh): the entries of one (lane, block column) are `base` (row k = 0) and Ml[0..6] (k = 1..7).
// Word j holds rows 2j and 2j+1 as bytes {A_2j, A_2j+1, B_2j, B_2j+1} = offset(2j) + 256 * offset(2j+1) on the packed pairs:
// one multiply-add per word on the FMA pipe (the ALU pipe is the one the recurrence saturates).
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint4 d8_encode(uint32_t base, const uint32_t (&Ml)[RPT])
{   // (Ml[2j] - base) * 256 + (Ml[2j-1] - base) = Ml[2j] * 256 + Ml[2j-1] - 257 * base in the ring of 32-bit words: nothing but
    // multiply-adds, none of them on the ALU pipe (seven packed subtractions per column before).  Neither half can borrow in
    // the result: the maxima grow with the row, every offset is >= 0.
    const uint32_t nb256 = base * 0xFFFFFF00u, nb257 = base * 0xFFFFFEFFu;          // -256 * base, -257 * base
    return make_uint4(mad_u32(Ml[0], 256u, nb256), mad_u32(Ml[1], 1u, mad_u32(Ml[2], 256u, nb257)),
                      mad_u32(Ml[3], 1u, mad_u32(Ml[4], 256u, nb257)), mad_u32(Ml[5], 1u, mad_u32(Ml[6], 256u, nb257)));
}
__device__ __forceinline__ uint32_t d8_offset(const uint4 &o, int k)       // packed offset of row k (0 in the upper bytes)
{
    const uint32_t wd = (k >> 1) == 0 ? o.x : (k >> 1) == 1 ? o.y : (k >> 1) == 2 ? o.z : o.w;
    return (k & 1) ? prmt(wd, 0u, 0x4341) : (wd & 0x00FF00FFu);
}

struct LaneState {
    uint32_t Hl[RPT], Gl[RPT], Ml[RPT];    // row state at the last finished column
    uint32_t oH[W], oG[W], oM[W];          // bottom-row outputs of the last finished block (shuffled to lane p+1)
    uint32_t pH, pM;                       // H / M of the row above at the last finished column (diagonal / D row shift)
};

struct WarpRings {                          // shared memory of one warp
    uint4 in[2][3][32];                     // double-buffered boundary feed of lane p = 0: planes H, G, M
    uint4 out[3][32];                       // boundary rows produced by lane p = 31
};

// Reverse main sweep: the D slots of the coming steps, staged in shared memory.  A slot is one contiguous piece of the
// diagonal-major plane (2560 B in the OFF8 format, 4096 B raw).  Bulk form: one elected lane issues a single bulk copy per
// step (cp.async.bulk, the TMA engine) that completes on the buffer's mbarrier; every lane then waits on the barrier's
// parity.  The 8 KB hold three OFF8 slots (fetched two steps ahead: a bulk copy takes longer to land than per-lane
// cp.async, and one step of lead leaves the warps spinning on the barrier) or two raw ones (one step ahead).
struct DStage {
    uint4 d[512];                           // NB slots of DSTEP uint4 each: [slot][plane][lane]
    uint64_t bar[3];                        // one mbarrier per slot
    uint32_t phase;                         // bit b: the parity slot b completes next (kept in a register while a pair is swept)
    uint32_t pad_;
};
__host__ __device__ constexpr int dstage_slots(int dfmt) { return (AGE_DSTAGE_BULK && dfmt == DFMT_OFF8) ? 3 : 2; }

__device__ __forceinline__ void dstage_init(DStage &DS, int lane)
{
#if AGE_DSTAGE_BULK
    if (lane == 0) { mbar_init(&DS.bar[0], 1); mbar_init(&DS.bar[1], 1); mbar_init(&DS.bar[2], 1); DS.phase = 0; mbar_init_fence(); }
    __syncwarp();
#endif
}

struct TraceOut {                           // trace re-sweep outputs (one half of the pair, one kind of trace): pages of ONE window
    uint2 *P;                               // 2-bit codes: [step in window][lane], 16 bits per block column, 2 bits per row
    uint32_t *hot;                          // MODE_TRACE_M, reverse: [step in window][lane], bit 8w+k = the cell holds Fmax + Rmax == max
    uint32_t *Mp;                           // MODE_TRACE_M, reverse, optional: the running maxima themselves (K7 rows)
    int sh;                                 // 0 / 16: the half whose trace is wanted
    int max2;
};

template <int DIR> __device__ __forceinline__ uint32_t shfl_prev(uint32_t v)
{
    return DIR > 0 ? __shfl_up_sync(0xFFFFFFFFu, v, 1) : __shfl_down_sync(0xFFFFFFFFu, v, 1);
}

// Everything one (pair, direction, strip) sweep needs, hoisted out of the PairTask once.
template <int DIR> struct StripCtx {
    int len1, nb, nsteps, nwin, s, q0, lane, p;
    size_t sbase;                           // s * nsteps
    uint32_t c0x, kabs, flip, gborder, m2, mm2;
    uint32_t Ta[RPT], Tb[RPT];
    uint32_t rmask[RPT];                    // wide arithmetic only: 0x1FFFF on rows of the contig, 0 on padding rows (Ar<true>::store)
    const uint4 *xblk;
    const uint8_t *ycode;
    const uint4 *bin; uint4 *bout;          // boundary planes in / out (nullptr at the first / last strip)
    uint4 *D4;                              // D as uint4
    int dfmt, dstep;                        // storage format of D and its uint4 per step (32 lanes x d_planes)
    uint4 *Dtail4;                          // forward, last lane of the last strip only (else nullptr)
    uint32_t *tilemax;
    uint32_t *ckpt;
    bool first;                             // no boundary input: the strip starts at the matrix border
    bool live;                              // the lane's rows of D (q0 .. q0+7) include rows of the matrix: padding lanes skip the D traffic
    volatile int *prog_out;                 // multi-warp sweep: blocks of this strip's boundary row flushed so far (or null)

    __device__ __forceinline__ void init(const PairTask &T, int s_, int si, int lane_)
    {
        constexpr bool FWD = DIR > 0;
        len1 = T.len1; nb = T.nb; nsteps = T.nsteps; nwin = T.nwin; s = s_; lane = lane_;
        p = FWD ? lane : 31 - lane;
        q0 = s * STRIP + lane * RPT;
        sbase = (size_t)s * nsteps;
        c0x = T.c0x; kabs = T.kabs; flip = T.flip; gborder = T.gborder; m2 = T.m2; mm2 = T.mm2;
#pragma unroll
        for (int k = 0; k < RPT; ++k) { const uint2 y = __ldg(T.ytab + q0 + k); Ta[k] = y.x; Tb[k] = y.y; rmask[k] = q0 + k < T.len2 ? 0x1FFFFu : 0u; }
        xblk = FWD ? T.xblk_f : T.xblk_r;
        ycode = T.ycode;
        uint4 *bnd = FWD ? T.bnd_f : T.bnd_r;
        first = si == 0;
        bin = first ? nullptr : bnd + (size_t)(si - 1) * 3 * (nb + 1);
        bout = (si + 1 < T.nstrips) ? bnd + (size_t)si * 3 * (nb + 1) : nullptr;
        D4 = reinterpret_cast<uint4 *>(T.D);
        dfmt = T.dfmt; dstep = 32 * d_planes(T.dfmt);
        Dtail4 = (FWD && p == 31 && si + 1 == T.nstrips) ? reinterpret_cast<uint4 *>(T.Dtail) : nullptr;
        tilemax = T.tilemax;
        ckpt = FWD ? T.ckpt_f : T.ckpt_r;
        prog_out = nullptr;
        live = q0 <= T.len2;
    }
};

// wide arithmetic: all-ones when column w of the lane's block b is a column of the window, 0 when it is padding behind len1
template <int DIR> __device__ __forceinline__ uint32_t col_mask(const StripCtx<DIR> &C, int b, int w)
{
    const int c = DIR > 0 ? W * (b - 1) + 1 + w : W * (C.nb - b + 1) - w;
    return c <= C.len1 ? 0xFFFFFFFFu : 0u;
}

__device__ __forceinline__ void state_init(LaneState &S, uint32_t gborder)
{
#pragma unroll
    for (int k = 0; k < RPT; ++k) { S.Hl[k] = 0; S.Gl[k] = gborder; S.Ml[k] = 0; }
#pragma unroll
    for (int w = 0; w < W; ++w) { S.oH[w] = 0; S.oG[w] = gborder; S.oM[w] = 0; }
    S.pH = 0; S.pM = 0;
}

__device__ __forceinline__ void state_save(const LaneState &S, uint32_t *dst /* + lane */)
{
    int i = 0;
#pragma unroll
    for (int k = 0; k < RPT; ++k) { dst[32 * i++] = S.Hl[k]; }
#pragma unroll
    for (int k = 0; k < RPT; ++k) { dst[32 * i++] = S.Gl[k]; }
#pragma unroll
    for (int k = 0; k < RPT; ++k) { dst[32 * i++] = S.Ml[k]; }
#pragma unroll
    for (int w = 0; w < W; ++w) { dst[32 * i++] = S.oH[w]; }
#pragma unroll
    for (int w = 0; w < W; ++w) { dst[32 * i++] = S.oG[w]; }
#pragma unroll
    for (int w = 0; w < W; ++w) { dst[32 * i++] = S.oM[w]; }
    dst[32 * i++] = S.pH; dst[32 * i++] = S.pM;
}

__device__ __forceinline__ void state_load(LaneState &S, const uint32_t *src /* + lane */)
{
    int i = 0;
#pragma unroll
    for (int k = 0; k < RPT; ++k) { S.Hl[k] = __ldcg(src + 32 * i++); }
#pragma unroll
    for (int k = 0; k < RPT; ++k) { S.Gl[k] = __ldcg(src + 32 * i++); }
#pragma unroll
    for (int k = 0; k < RPT; ++k) { S.Ml[k] = __ldcg(src + 32 * i++); }
#pragma unroll
    for (int w = 0; w < W; ++w) { S.oH[w] = __ldcg(src + 32 * i++); }
#pragma unroll
    for (int w = 0; w < W; ++w) { S.oG[w] = __ldcg(src + 32 * i++); }
#pragma unroll
    for (int w = 0; w < W; ++w) { S.oM[w] = __ldcg(src + 32 * i++); }
    S.pH = __ldcg(src + 32 * i++); S.pM = __ldcg(src + 32 * i++);
}

// asynchronous copy of the boundary blocks that lane p = 0 consumes in window `win` into ring buffer `buf`
template <int DIR>
__device__ __forceinline__ void ring_prefetch(const StripCtx<DIR> &C, WarpRings &R, int win, int buf)
{
    const int b = CKS * win + 1 + C.lane;
    if (C.bin && b <= C.nb) {
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) cp_async16(&R.in[buf][pl][C.lane], C.bin + (size_t)pl * (C.nb + 1) + b);
    }
    cp_async_commit();
}

__device__ __forceinline__ void ring_fill_border(WarpRings &R, int lane, uint32_t gborder)
{
#pragma unroll
    for (int buf = 0; buf < 2; ++buf) {
        R.in[buf][0][lane] = make_uint4(0, 0, 0, 0);
        R.in[buf][1][lane] = make_uint4(gborder, gborder, gborder, gborder);
        R.in[buf][2][lane] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
}

__device__ __forceinline__ void stcg16(uint4 *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{   // the D stream is written once and read once, much later: keep it out of L1
    asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// D slot of reverse step t (written by the forward sweep in step nb+32-t) -> staging slot t mod NB, asynchronously
template <int DIR>
__device__ __forceinline__ void dstage_fetch(const StripCtx<DIR> &C, DStage &DS, int t)
{
    const int b = t % dstage_slots(C.dfmt);
#if AGE_DSTAGE_BULK
    if (t <= C.nsteps && C.lane == 0) {     // (lanes whose rows lie beyond the contig get their part of the slot too: never looked at)
        const uint32_t bytes = (uint32_t)C.dstep * 16u;
        mbar_expect_tx(&DS.bar[b], bytes);
        bulk_g2s(DS.d + b * C.dstep, C.D4 + (C.sbase + (size_t)(C.nb + 31 - t)) * C.dstep, bytes, &DS.bar[b]);
    }
#else
    if (t <= C.nsteps && C.live) {
        const uint4 *src = C.D4 + (C.sbase + (size_t)(C.nb + 31 - t)) * C.dstep + C.lane;
#pragma unroll
        for (int i = 0; i < W * 2; ++i) if (i < 5 || C.dfmt == DFMT_RAW) cp_async16(DS.d + b * C.dstep + i * 32 + C.lane, src + i * 32);
    }
    cp_async_commit();
#endif
}
// the D slot of step t (staging slot b) has landed
__device__ __forceinline__ void dstage_wait(DStage &DS, int b, uint32_t &dph)
{
#if AGE_DSTAGE_BULK
    mbar_wait(&DS.bar[b], (dph >> b) & 1u);
    dph ^= 1u << b;
#else
    asm volatile("cp.async.wait_group 1;" ::: "memory");
#endif
}

// One window (steps t0 .. t1) of one strip of the main sweep in which some lanes are idle in some steps: the first
// window(s) of a strip (the wavefront fills up), the last ones (it drains, and the last block may hold padding columns).
// The windows in between go through sweep_window_steady.
template <int DIR, int DFMT, bool WIDE>
__device__ __forceinline__ void sweep_window_edge(const StripCtx<DIR> &C, LaneState &S, WarpRings &R, DStage &DS, const int win,
                                                  const int t0, const int t1, uint32_t &best, uint32_t &dph)
{
    using A = Ar<WIDE>;
    constexpr bool FWD = DIR > 0;
    constexpr int DSTEP = 32 * (DFMT == DFMT_OFF8 ? 5 : 2 * W), NB = dstage_slots(DFMT);
    const int lane = C.lane, p = C.p, nb = C.nb;
    const int buf = win & 1;
    auto xload = [&](int b) { return __ldg(C.xblk + min(max(b, 0), nb + 1)); };
    uint4 xnext = xload(t0 - p);

    for (int t = t0; t <= t1; ++t) {
        const uint4 xb = xnext;
        xnext = xload(t + 1 - p);
        // ---- values of the row above for the 4 columns of this block -------------------------------------
        uint32_t uH[W], uG[W], uM[W];
#pragma unroll
        for (int w = 0; w < W; ++w) { uH[w] = shfl_prev<DIR>(S.oH[w]); uG[w] = shfl_prev<DIR>(S.oG[w]); uM[w] = shfl_prev<DIR>(S.oM[w]); }
        const uint4 *dcur = DS.d + (t % NB) * DSTEP;
        if (!FWD) {                             // D of the coming step(s) in flight, D of this step landed (each lane reads its own).
            dstage_fetch(C, DS, t + NB - 1);    // Behind the shuffles: every lane is through with the slot that is refilled.
            dstage_wait(DS, t % NB, dph);
        }
        if (p == 0) {
            const int e = (t - 1) & 31;
            const uint4 h = R.in[buf][0][e], g = R.in[buf][1][e], m = R.in[buf][2][e];
            uH[0] = h.x; uH[1] = h.y; uH[2] = h.z; uH[3] = h.w;
            uG[0] = g.x; uG[1] = g.y; uG[2] = g.z; uG[3] = g.w;
            uM[0] = m.x; uM[1] = m.y; uM[2] = m.z; uM[3] = m.w;
        }
        const int b = t - p;
        if (b >= 1 && b <= nb) {
            uint4 *dst = C.D4 + (C.sbase + (size_t)(t - 1)) * DSTEP + lane;      // forward: the D slot of this step
            if (FWD && C.live && DFMT == DFMT_OFF8) stcg16(dst, S.pM, uM[0], uM[1], uM[2]);
            uint4 dbase = make_uint4(0, 0, 0, 0);
            if (!FWD && C.live && DFMT == DFMT_OFF8) dbase = dcur[lane];
            const bool special = (xb.z & 0x01010101u) != 0;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                // ---- forward: D = maxima at the previous column, rows shifted by one (entry 0 = row above) ----------
                const uint32_t upMprev = w == 0 ? S.pM : uM[w - 1];
                if (FWD && C.live) {
                    if (DFMT == DFMT_RAW) {
                        stcg16(dst + w * 64, upMprev, S.Ml[0], S.Ml[1], S.Ml[2]);
                        stcg16(dst + w * 64 + 32, S.Ml[3], S.Ml[4], S.Ml[5], S.Ml[6]);
                    } else {
                        const uint4 o = d8_encode(upMprev, S.Ml);
                        stcg16(dst + (1 + w) * 32, o.x, o.y, o.z, o.w);
                    }
                }
                // ---- substitution scores of this column ---------------------------------------------------------
                uint32_t S2[RPT];
                const uint32_t sel = ((w < 2 ? xb.x : xb.y) >> (16 * (w & 1))) & 0xFFFFu;
                const uint32_t info = (xb.z >> (8 * w)) & 0xFFu;
                if (special && (info & 1u)) {
                    const int xa = (info >> 1) & 7, xc = (info >> 4) & 7;
                    const int m2a = A::lo(C.m2), m2b = A::hi(C.m2), mm2a = A::lo(C.mm2), mm2b = A::hi(C.mm2);
#pragma unroll
                    for (int k = 0; k < RPT; ++k) {
                        const int yc = C.ycode[C.q0 + k];
                        const int sa = subst2(xa, yc & 15, m2a, mm2a), sb = subst2(xc, yc >> 4, m2b, mm2b);
                        S2[k] = A::pack(sa, sb);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < RPT; ++k) S2[k] = prmt(C.Ta[k], C.Tb[k], sel);
                }
                // ---- the column ---------------------------------------------------------------------------------
                uint4 dc0 = make_uint4(0, 0, 0, 0), dc1 = dc0;
                uint32_t colbase = 0, colbest = 0;
                if (!FWD && C.live) {
                    if (DFMT == DFMT_RAW) { dc0 = dcur[((W - 1 - w) * 2 + 0) * 32 + lane]; dc1 = dcur[((W - 1 - w) * 2 + 1) * 32 + lane]; }
                    else {
                        dc0 = dcur[(1 + (W - 1 - w)) * 32 + lane];
                        colbase = (W - 1 - w) == 0 ? dbase.x : (W - 1 - w) == 1 ? dbase.y : (W - 1 - w) == 2 ? dbase.z : dbase.w;
                    }
                }
                const uint32_t cmask = WIDE ? col_mask(C, b, w) : 0xFFFFFFFFu;
                uint32_t vG = uG[w], vM = uM[w], dg = w == 0 ? S.pH : uH[w - 1];
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    const int k = FWD ? j : RPT - 1 - j;
                    const uint32_t g  = A::maxs(S.Gl[k], vG);
                    const uint32_t s2 = A::store(A::addmax_relu(dg, S2[k], g), C.rmask[k], cmask);    // max(2d, 2h+1, 2v+1, 0)
                    const uint32_t tg = (s2 ^ C.flip) & A::TAG;
                    const uint32_t G2 = A::add(tg * C.kabs + s2, C.c0x);          // 2*(s + (gap ? ge : go)) + 1
                    const uint32_t Hc = s2 & A::NTAG;
                    const uint32_t Mn = A::max3(Hc, S.Ml[k], vM);
                    if (!FWD) {
                        if (DFMT == DFMT_RAW) {
                            const uint4 dv = (k >> 2) ? dc1 : dc0;
                            const uint32_t dk = (k & 3) == 0 ? dv.x : (k & 3) == 1 ? dv.y : (k & 3) == 2 ? dv.z : dv.w;
                            best = A::maxu(best, A::sum(A::add(Mn, dk)));
                        } else colbest = __vmaxu2(colbest, __vadd2(Mn, d8_offset(dc0, k)));
                    }
                    dg = S.Hl[k];
                    S.Hl[k] = Hc; S.Gl[k] = G2; S.Ml[k] = Mn;
                    vG = G2; vM = Mn;
                }
                if (!FWD && DFMT == DFMT_OFF8) best = __vmaxu2(best, __vadd2(colbest, colbase));
                constexpr int last = FWD ? RPT - 1 : 0;
                S.oH[w] = S.Hl[last]; S.oG[w] = S.Gl[last]; S.oM[w] = S.Ml[last];
            }
            S.pH = uH[W - 1]; S.pM = uM[W - 1];
            if (FWD && C.Dtail4) C.Dtail4[b - 1] = make_uint4(S.oM[0], S.oM[1], S.oM[2], S.oM[3]);
            if (p == 31 && C.bout) {
                const int e = (b - 1) & 31;
                R.out[0][e] = make_uint4(S.oH[0], S.oH[1], S.oH[2], S.oH[3]);
                R.out[1][e] = make_uint4(S.oG[0], S.oG[1], S.oG[2], S.oG[3]);
                R.out[2][e] = make_uint4(S.oM[0], S.oM[1], S.oM[2], S.oM[3]);
            }
        }
        if (C.bout) {
            const int b31 = t - 31;                      // block just finished by the last logical lane
            if (b31 >= 1 && ((b31 & 31) == 0 || b31 == nb)) {
                __syncwarp();
                const int bb = ((b31 - 1) & ~31) + 1 + lane;
                if (bb <= b31) {
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl) C.bout[(size_t)pl * (nb + 1) + bb] = R.out[pl][lane];
                }
                if (C.prog_out) { __threadfence(); __syncwarp(); if (lane == 0) *C.prog_out = b31; }   // the next strip's warp may read them
                __syncwarp();
            }
        }
    }
}

// The steady-state window of the main sweep (every lane active in every step, 32 <= t0, t1 <= nb), written for
// instruction count -- running pointers instead of per-step address arithmetic,
// no clamps on the column descriptors (1 <= t - p <= nb throughout), and the N / IUPAC test hoisted out of the
// column loop (one copy of the step body without it, one with it).
template <int DIR, int DFMT, bool FLIP, bool WIDE>
__device__ __forceinline__ void sweep_window_steady(const StripCtx<DIR> &C, LaneState &S, WarpRings &R, DStage &DS, const int win,
                                                    const int t0, const int t1, uint32_t &best, uint32_t &dph)
{
    using A = Ar<WIDE>;
    constexpr bool FWD = DIR > 0;
    constexpr int NPL = DFMT == DFMT_OFF8 ? 5 : 2 * W, DSTEP = 32 * NPL;     // uint4 of D per (lane, step) / per step
    constexpr int NB = dstage_slots(DFMT);                                    // staging slots; the fetch runs NB - 1 steps ahead
    const int lane = C.lane, p = C.p, nb = C.nb;
    const int buf = win & 1;
    // Column descriptors: only the selectors (.x, .y) are loaded per step, one step ahead.  The per-step "has an N / IUPAC
    // column" branch must NOT depend on that load: ptxas schedules the test right behind the LDG (to free its register) and
    // every step then waits a full memory latency (37 % of all stall samples of the kernel before this change).  The flags
    // of the lane's 32 blocks of this window come from the .w words instead (see age_pack_kernel), once per window.
    const uint4 *xp = C.xblk + (t0 - p);
    uint32_t spec;
    {
        const uint32_t a = __ldg(&xp[0].w), b = __ldg(&xp[t1 - t0].w);
        spec = __funnelshift_r(a, b, (t0 - p) & 31);
        if (t1 - t0 < 31) spec &= (1u << (t1 - t0 + 1)) - 1u;
    }
    uint2 xnext = __ldg(reinterpret_cast<const uint2 *>(xp));
    // forward: D slot of step t; reverse: D slot of the step fetched next (t + NB - 1), i.e. forward step nb + 32 - (t + NB - 1)
    uint4 *dptr = C.D4 + (C.sbase + (size_t)(FWD ? t0 - 1 : nb + 32 - NB - t0)) * DSTEP + lane;
    const bool live = C.live;
    int cb = t0 % NB;                                // staging slot of step t

    for (int t = t0; t <= t1; ++t) {
        const uint2 xb = xnext;
        xnext = __ldg(reinterpret_cast<const uint2 *>(++xp));
        const bool special = spec & 1u;
        spec >>= 1;
        uint32_t uH[W], uG[W], uM[W];
#pragma unroll
        for (int w = 0; w < W; ++w) { uH[w] = shfl_prev<DIR>(S.oH[w]); uG[w] = shfl_prev<DIR>(S.oG[w]); uM[w] = shfl_prev<DIR>(S.oM[w]); }
        const uint4 *dcur = DS.d + cb * DSTEP;
        if (!FWD) {                             // behind the shuffles: every lane is through with the slot that is refilled
            const int pb = cb == 0 ? NB - 1 : cb - 1;        // slot of step t + NB - 1 (<= nb + 2 <= nsteps: always a valid step here)
#if AGE_DSTAGE_BULK
            if (lane == 0) {
                mbar_expect_tx(&DS.bar[pb], DSTEP * 16);
                bulk_g2s(DS.d + pb * DSTEP, dptr - lane, DSTEP * 16, &DS.bar[pb]);
            }
#else
            if (live) {
#pragma unroll
                for (int i = 0; i < NPL; ++i) cp_async16(DS.d + pb * DSTEP + i * 32 + lane, dptr + i * 32);
            }
            cp_async_commit();
#endif
            dptr -= DSTEP;
            dstage_wait(DS, cb, dph);
        }
        if (p == 0) {
            const int e = (t - 1) & 31;
            const uint4 h = R.in[buf][0][e], g = R.in[buf][1][e], m = R.in[buf][2][e];
            uH[0] = h.x; uH[1] = h.y; uH[2] = h.z; uH[3] = h.w;
            uG[0] = g.x; uG[1] = g.y; uG[2] = g.z; uG[3] = g.w;
            uM[0] = m.x; uM[1] = m.y; uM[2] = m.z; uM[3] = m.w;
        }
        if (FWD && live && DFMT == DFMT_OFF8) stcg16(dptr, S.pM, uM[0], uM[1], uM[2]);       // entry k = 0 of the four columns
        uint4 dbase = make_uint4(0, 0, 0, 0);
        if (!FWD && live && DFMT == DFMT_OFF8) dbase = dcur[lane];
        auto columns = [&](auto sp) {
            constexpr bool SPECIAL = decltype(sp)::value;
            // the per-column N / IUPAC codes are only fetched on the rare path: even a predicated-off load in the common
            // path costs a scoreboard wait on its destination register
            const uint32_t xz = SPECIAL ? __ldg(&xp[-1].z) : 0u;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                if (FWD && live) {
                    const uint32_t upMprev = w == 0 ? S.pM : uM[w - 1];
                    if (DFMT == DFMT_RAW) {
                        stcg16(dptr + w * 64, upMprev, S.Ml[0], S.Ml[1], S.Ml[2]);
                        stcg16(dptr + w * 64 + 32, S.Ml[3], S.Ml[4], S.Ml[5], S.Ml[6]);
                    } else {
                        const uint4 o = d8_encode(upMprev, S.Ml);
                        stcg16(dptr + (1 + w) * 32, o.x, o.y, o.z, o.w);
                    }
                }
                uint32_t S2[RPT];
                const uint32_t sel = ((w < 2 ? xb.x : xb.y) >> (16 * (w & 1))) & 0xFFFFu;
                const uint32_t info = SPECIAL ? (xz >> (8 * w)) & 0xFFu : 0u;
                if (SPECIAL && (info & 1u)) {
                    const int xa = (info >> 1) & 7, xc = (info >> 4) & 7;
                    const int m2a = A::lo(C.m2), m2b = A::hi(C.m2), mm2a = A::lo(C.mm2), mm2b = A::hi(C.mm2);
#pragma unroll
                    for (int k = 0; k < RPT; ++k) {
                        const int yc = C.ycode[C.q0 + k];
                        const int sa = subst2(xa, yc & 15, m2a, mm2a), sb = subst2(xc, yc >> 4, m2b, mm2b);
                        S2[k] = A::pack(sa, sb);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < RPT; ++k) S2[k] = prmt(C.Ta[k], C.Tb[k], sel);
                }
                uint4 dc0 = make_uint4(0, 0, 0, 0), dc1 = dc0;
                uint32_t colbase = 0, colbest = 0;
                if (!FWD && live) {
                    if (DFMT == DFMT_RAW) { dc0 = dcur[((W - 1 - w) * 2 + 0) * 32 + lane]; dc1 = dcur[((W - 1 - w) * 2 + 1) * 32 + lane]; }
                    else {
                        dc0 = dcur[(1 + (W - 1 - w)) * 32 + lane];
                        colbase = (W - 1 - w) == 0 ? dbase.x : (W - 1 - w) == 1 ? dbase.y : (W - 1 - w) == 2 ? dbase.z : dbase.w;
                    }
                }
                const uint32_t cmask = WIDE ? col_mask(C, t - p, w) : 0xFFFFFFFFu;
                uint32_t vG = uG[w], vM = uM[w], dg = w == 0 ? S.pH : uH[w - 1];
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    const int k = FWD ? j : RPT - 1 - j;
                    const uint32_t g  = A::maxs(S.Gl[k], vG);
                    const uint32_t s2 = A::store(A::addmax_relu(dg, S2[k], g), C.rmask[k], cmask);
                    const uint32_t tg = FLIP ? (~s2 & A::TAG) : (s2 & A::TAG);
                    const uint32_t G2 = A::add(tg * C.kabs + s2, C.c0x);
                    // clearing the tags: with FLIP = false tg IS the tag pair, and s2 - tg cannot borrow between the halves,
                    // so the subtraction can run as a multiply-add on the FMA pipe (the ALU pipe is the busy one)
                    uint32_t Hc;
                    if (FLIP) Hc = s2 & A::NTAG;
                    else asm("mad.lo.u32 %0, %1, 0xFFFFFFFF, %2;" : "=r"(Hc) : "r"(tg), "r"(s2));
                    const uint32_t Mn = A::max3(Hc, S.Ml[k], vM);
                    if (!FWD) {
                        if (DFMT == DFMT_RAW) {
                            const uint4 dv = (k >> 2) ? dc1 : dc0;
                            const uint32_t dk = (k & 3) == 0 ? dv.x : (k & 3) == 1 ? dv.y : (k & 3) == 2 ? dv.z : dv.w;
                            best = A::maxu(best, A::sum(A::add(Mn, dk)));
                        } else colbest = __vmaxu2(colbest, __vadd2(Mn, d8_offset(dc0, k)));
                    }
                    dg = S.Hl[k];
                    S.Hl[k] = Hc; S.Gl[k] = G2; S.Ml[k] = Mn;
                    vG = G2; vM = Mn;
                }
                if (!FWD && DFMT == DFMT_OFF8) best = __vmaxu2(best, __vadd2(colbest, colbase));
                constexpr int last = FWD ? RPT - 1 : 0;
                S.oH[w] = S.Hl[last]; S.oG[w] = S.Gl[last]; S.oM[w] = S.Ml[last];
            }
        };
        if (!special) columns(std::false_type{});
        else                           columns(std::true_type{});
        if (FWD) dptr += DSTEP;
        cb = cb + 1 == NB ? 0 : cb + 1;
        S.pH = uH[W - 1]; S.pM = uM[W - 1];
        const int b = t - p;
        if (FWD && C.Dtail4) C.Dtail4[b - 1] = make_uint4(S.oM[0], S.oM[1], S.oM[2], S.oM[3]);
        if (C.bout) {
            if (p == 31) {
                const int e = (b - 1) & 31;
                R.out[0][e] = make_uint4(S.oH[0], S.oH[1], S.oH[2], S.oH[3]);
                R.out[1][e] = make_uint4(S.oG[0], S.oG[1], S.oG[2], S.oG[3]);
                R.out[2][e] = make_uint4(S.oM[0], S.oM[1], S.oM[2], S.oM[3]);
            }
            const int b31 = t - 31;                      // block just finished by the last logical lane
            if (b31 >= 1 && ((b31 & 31) == 0 || b31 == nb)) {
                __syncwarp();
                const int bb = ((b31 - 1) & ~31) + 1 + lane;
                if (bb <= b31) {
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl) C.bout[(size_t)pl * (nb + 1) + bb] = R.out[pl][lane];
                }
                if (C.prog_out) { __threadfence(); __syncwarp(); if (lane == 0) *C.prog_out = b31; }
                __syncwarp();
            }
        }
    }
}

// dispatch on whether the window is steady (all lanes active in all of its steps)
template <int DIR, int DFMT, bool WIDE>
__device__ __forceinline__ void sweep_window_any(const StripCtx<DIR> &C, LaneState &S, WarpRings &R, DStage &DS, int win, uint32_t &best, uint32_t &dph)
{
    const int t0 = CKS * win + 1, t1 = min(CKS * win + CKS, C.nsteps);
    if (t0 >= 32 && t1 <= C.nb) {
        if (C.flip == 0) sweep_window_steady<DIR, DFMT, false, WIDE>(C, S, R, DS, win, t0, t1, best, dph);
        else             sweep_window_steady<DIR, DFMT, true, WIDE>(C, S, R, DS, win, t0, t1, best, dph);
    }
    else                        sweep_window_edge<DIR, DFMT, WIDE>(C, S, R, DS, win, t0, t1, best, dph);
}

// The main sweep of one pair in one direction: checkpoints, tile maxima, all windows of the strips si0, si0 + stride, ...
// (in sweep order).  With stride > 1 several warps of one CTA share the pair (long alignments): strip si starts as soon
// as the warp of strip si-1 has flushed the first boundary blocks and then runs 32+ steps behind it; `prog` (shared
// memory, one counter per strip, zeroed by the caller) carries the number of flushed blocks.
template <int DIR, bool WIDE>
__device__ void sweep_pair_fast(const PairTask &T, WarpRings &R, DStage &DS, const int lane,
                                const int si0 = 0, const int stride = 1, volatile int *prog = nullptr)
{
    constexpr bool FWD = DIR > 0;
    uint32_t dph = FWD ? 0u : DS.phase;          // parities of the D staging barriers, in a register while the pair is swept
    for (int si = si0; si < T.nstrips; si += stride) {
        const int s = FWD ? si : T.nstrips - 1 - si;
        StripCtx<DIR> C;
        C.init(T, s, si, lane);
        if (prog) C.prog_out = prog + si;
        volatile int *prog_in = (prog && !C.first) ? prog + si - 1 : nullptr;
        auto wait_blocks = [&](int win) {       // boundary blocks of window `win` must have been flushed by the producer
            if (!prog_in) return;
            const int need = min(CKS * win + CKS, C.nb);
            while (*prog_in < need) __nanosleep(200);
            __threadfence();
        };
        LaneState S;
        state_init(S, C.gborder);
        if (C.first) ring_fill_border(R, lane, C.gborder);
        else { wait_blocks(0); ring_prefetch(C, R, 0, 0); }
        uint32_t best = 0;
        if (!FWD) for (int t = 1; t < dstage_slots(C.dfmt); ++t) dstage_fetch(C, DS, t);
        for (int win = 0; win < C.nwin; ++win) {
            cp_async_wait_all();
            __syncwarp();
            if (!C.first && win + 1 < C.nwin) { wait_blocks(win + 1); ring_prefetch(C, R, win + 1, (win + 1) & 1); }
            state_save(S, C.ckpt + ((size_t)(s * C.nwin + win) * CKW) * 32 + lane);
            if (!WIDE && C.dfmt == DFMT_OFF8) sweep_window_any<DIR, DFMT_OFF8, false>(C, S, R, DS, win, best, dph);
            else                              sweep_window_any<DIR, DFMT_RAW, WIDE>(C, S, R, DS, win, best, dph);
            if (!FWD) { C.tilemax[(size_t)(s * C.nwin + win) * 32 + lane] = best; best = 0; }
        }
        // the maxima at the last processed column: Fmax[.][len1] (forward) / Rmax[.][1] (reverse)
        uint32_t *col = FWD ? T.Dlast : T.Rfirst;
#pragma unroll
        for (int k = 0; k < RPT; ++k) col[C.q0 + k + 1] = S.Ml[k];
        if (s == 0 && lane == 0) col[0] = 0;
        __syncwarp();
    }
    if (!FWD) { if (lane == 0) DS.phase = dph; __syncwarp(); }
}

// Shared-memory staging of the compact trace re-sweep (one per warp): the per-column values that the fast sweep
// keeps in statically indexed registers.
struct TraceSmem {
    uint32_t uH[W][32], uG[W][32], uM[W][32];   // values of the row above, per block column
    uint32_t oH[W][32], oG[W][32], oM[W][32];   // bottom-row outputs of the lane's last finished block
    uint32_t tw[W][32];                         // 2-bit codes of the 8 rows, per block column
};

// The trace re-sweep of pass 2: the same recurrence as sweep_window, written compactly -- the loop over the four
// columns of a block is NOT unrolled (its per-column values live in shared memory), so the whole step body is a few
// hundred instructions.  Pass 2 is latency- and instruction-fetch-bound: a fully unrolled 32-cell body with the
// trace code inlined overflows the instruction cache and runs several times slower.
template <int DIR, int MODE, bool WIDE>
__device__ __forceinline__ void sweep_window_compact(const StripCtx<DIR> &C, LaneState &S, WarpRings &R, TraceSmem &X,
                                                  const int win, const TraceOut &TO)
{
    using A = Ar<WIDE>;
    constexpr bool FWD = DIR > 0;
    constexpr bool TRACE_S = MODE == MODE_TRACE_S, TRACE_M = MODE == MODE_TRACE_M;
    const int lane = C.lane, p = C.p, nb = C.nb, buf = win & 1;
    const int t0 = CKS * win + 1, t1 = min(CKS * win + CKS, C.nsteps);
    const int prev_lane = FWD ? (lane + 31) & 31 : (lane + 1) & 31;
    // the wanted half of a packed word (wide: the whole word)
    const uint32_t hbit = 1u << TO.sh, hmask = WIDE ? 0xFFFFFFFFu : 0xFFFFu << TO.sh, hsign = WIDE ? 0x80000000u : 0x8000u << TO.sh;
    const uint32_t maxw = (uint32_t)TO.max2 << TO.sh;
#pragma unroll
    for (int w = 0; w < W; ++w) { X.oH[w][lane] = S.oH[w]; X.oG[w][lane] = S.oG[w]; X.oM[w][lane] = S.oM[w]; }

    // Loads that would otherwise sit in the step's dependency chain run ahead of it (AGE_V_PREFETCH): the column descriptors by
    // one step, and (reverse maxima trace) the forward maxima D by one block column -- two dependent DRAM latencies per column else.
    uint4 xnext = __ldg(C.xblk + min(max(t0 - p, 0), nb + 1));
    uint4 pa = make_uint4(0, 0, 0, 0), pb = pa;
    auto dfetch = [&](int tt, int ww, uint4 &da, uint4 &db) {      // D of (step tt, block column ww)
        da = make_uint4(0, 0, 0, 0); db = da;
        if (!FWD && TRACE_M && C.live) {
            const int bb = tt - p;
            if (tt <= t1 && bb >= 1 && bb <= nb) {
                const uint4 *slot = C.D4 + (C.sbase + (size_t)(nb + 31 - tt)) * C.dstep + lane;
                if (C.dfmt == DFMT_RAW) { da = __ldcg(slot + (W - 1 - ww) * 64); db = __ldcg(slot + (W - 1 - ww) * 64 + 32); }
                else { db.x = __ldcg(reinterpret_cast<const uint32_t *>(slot) + (W - 1 - ww)); da = __ldcg(slot + (1 + (W - 1 - ww)) * 32); }
            }
        }
    };
    if (AGE_V_PREFETCH) dfetch(t0, 0, pa, pb);

    for (int t = t0; t <= t1; ++t) {
        const int b = t - p;
        const uint4 xb = AGE_V_PREFETCH ? xnext : __ldg(C.xblk + min(max(b, 0), nb + 1));
        if (AGE_V_PREFETCH) xnext = __ldg(C.xblk + min(max(b + 1, 0), nb + 1));
        __syncwarp();
        uint32_t uH[W], uG[W], uM[W];
#pragma unroll
        for (int w = 0; w < W; ++w) { uH[w] = X.oH[w][prev_lane]; uG[w] = X.oG[w][prev_lane]; uM[w] = X.oM[w][prev_lane]; }
        if (p == 0) {
            const int e = (t - 1) & 31;
            const uint4 h = R.in[buf][0][e], g = R.in[buf][1][e], m = R.in[buf][2][e];
            uH[0] = h.x; uH[1] = h.y; uH[2] = h.z; uH[3] = h.w;
            uG[0] = g.x; uG[1] = g.y; uG[2] = g.z; uG[3] = g.w;
            uM[0] = m.x; uM[1] = m.y; uM[2] = m.z; uM[3] = m.w;
        }
#pragma unroll
        for (int w = 0; w < W; ++w) { X.uH[w][lane] = uH[w]; X.uG[w][lane] = uG[w]; X.uM[w][lane] = uM[w]; }
        __syncwarp();
        if (b < 1 || b > nb) { if (AGE_V_PREFETCH) dfetch(t + 1, 0, pa, pb); continue; }
        const int slot = t - t0;                  // step within the window: index into the window's pages
        uint32_t hotw = 0;
        uint32_t dg = S.pH;
#pragma unroll 1
        for (int w = 0; w < W; ++w) {
            const uint32_t upH = X.uH[w][lane], upG = X.uG[w][lane], upM = X.uM[w][lane];
            uint32_t S2[RPT];
            const uint32_t info = (xb.z >> (8 * w)) & 0xFFu;
            if (info & 1u) {
                const int xa = (info >> 1) & 7, xc = (info >> 4) & 7;
                const int m2a = A::lo(C.m2), m2b = A::hi(C.m2), mm2a = A::lo(C.mm2), mm2b = A::hi(C.mm2);
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const int yc = C.ycode[C.q0 + k];
                    const int sa = subst2(xa, yc & 15, m2a, mm2a), sb = subst2(xc, yc >> 4, m2b, mm2b);
                    S2[k] = A::pack(sa, sb);
                }
            } else {
                const uint32_t sel = ((w < 2 ? xb.x : xb.y) >> (16 * (w & 1))) & 0xFFFFu;
#pragma unroll
                for (int k = 0; k < RPT; ++k) S2[k] = prmt(C.Ta[k], C.Tb[k], sel);
            }
            uint32_t dk[RPT] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (!FWD && TRACE_M) {              // shifted forward maxima of this column (for the hot bits)
                uint4 ca = pa, cb = pb;
                if (AGE_V_PREFETCH) { if (w + 1 < W) dfetch(t, w + 1, pa, pb); else dfetch(t + 1, 0, pa, pb); }
                else dfetch(t, w, ca, cb);
                if (C.live) {
                    if (C.dfmt == DFMT_RAW) {
                        dk[0] = ca.x; dk[1] = ca.y; dk[2] = ca.z; dk[3] = ca.w; dk[4] = cb.x; dk[5] = cb.y; dk[6] = cb.z; dk[7] = cb.w;
                    } else {
#pragma unroll
                        for (int k = 0; k < RPT; ++k) dk[k] = cb.x + d8_offset(ca, k);
                    }
                }
            }
            const uint32_t cmask = WIDE ? col_mask(C, b, w) : 0xFFFFFFFFu;
            uint32_t vG = upG, vM = upM, codes = 0;
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const int k = FWD ? j : RPT - 1 - j;
                const uint32_t g  = A::maxs(S.Gl[k], vG);
                const uint32_t s2 = A::store(A::addmax_relu(dg, S2[k], g), C.rmask[k], cmask);
                const uint32_t tg = (s2 ^ C.flip) & A::TAG;
                const uint32_t G2 = A::add(tg * C.kabs + s2, C.c0x);
                const uint32_t Hc = s2 & A::NTAG;
                const uint32_t Mn = A::max3(Hc, S.Ml[k], vM);
                // Trace codes of the wanted half, read off the packed words the recurrence has anyway.
                //   FS (:1015-1017: s = d, t = D; if h >= s: H; if v >= s: V; if s < 0: stop): the cell's tag says gap (H / V)
                //       or not; among the gaps V wins iff the vertical candidate is the larger-or-equal one (g == vG);
                //       without a gap the cell is D unless even the diagonal candidate was negative.
                //   FM (:899-915 / :952-968: V if up > own, then H if left > the larger of the two): H iff the left maximum
                //       beats max(own, up), else V iff up beats own.
                if (TRACE_S) {
                    const bool gap = (s2 & hbit) != 0, vwin = ((g ^ vG) & hmask) == 0, dneg = (A::add(dg, S2[k]) & hsign) != 0;
                    codes |= (gap ? (vwin ? T_VERT : T_HORI) : (dneg ? T_STOP : T_DIAG)) << (2 * k);
                }
                if (TRACE_M) {
                    const uint32_t m1 = A::maxs(Hc, vM);
                    const bool isH = ((Mn ^ m1) & hmask) != 0, isV = ((m1 ^ Hc) & hmask) != 0;
                    codes |= (isH ? T_HORI : (isV ? T_VERT : T_STOP)) << (2 * k);
                }
                if (!FWD && TRACE_M) {
                    if (((A::add(Mn, dk[k]) ^ maxw) & hmask) == 0) hotw |= 1u << (8 * w + k);    // plain int compare (:565), no wrap
                }
                dg = S.Hl[k];
                S.Hl[k] = Hc; S.Gl[k] = G2; S.Ml[k] = Mn;
                vG = G2; vM = Mn;
            }
            constexpr int last = FWD ? RPT - 1 : 0;
            X.oH[w][lane] = S.Hl[last]; X.oG[w][lane] = S.Gl[last]; X.oM[w][lane] = S.Ml[last];
            X.tw[w][lane] = codes;
            if (TRACE_M && !FWD && TO.Mp) {
                uint4 *dst = reinterpret_cast<uint4 *>(TO.Mp) + (slot * W + w) * 64 + lane;      // (step, column) -> 2 x 32 uint4
                dst[0]  = make_uint4(S.Ml[0], S.Ml[1], S.Ml[2], S.Ml[3]);
                dst[32] = make_uint4(S.Ml[4], S.Ml[5], S.Ml[6], S.Ml[7]);
            }
            dg = upH;                           // diagonal of the next column's first row
        }
        S.pH = X.uH[W - 1][lane]; S.pM = X.uM[W - 1][lane];
        TO.P[slot * 32 + lane] = make_uint2(X.tw[0][lane] | (X.tw[1][lane] << 16), X.tw[2][lane] | (X.tw[3][lane] << 16));
        if (!FWD && TRACE_M) TO.hot[slot * 32 + lane] = hotw;
    }
    __syncwarp();
}

// Recompute one window of one strip from its checkpoint, emitting the trace (pass 2).
template <int DIR, int MODE, bool WIDE>
__device__ __forceinline__ void sweep_window_trace(const PairTask &T, WarpRings &R, TraceSmem &X, const int lane, int s, int win, const TraceOut &TO)
{
    constexpr bool FWD = DIR > 0;
    const int si = FWD ? s : T.nstrips - 1 - s;
    StripCtx<DIR> C;
    C.init(T, s, si, lane);
    LaneState S;
    state_load(S, C.ckpt + ((size_t)(s * C.nwin + win) * CKW) * 32 + lane);
    __syncwarp();
    if (C.first) ring_fill_border(R, lane, C.gborder);
    else { ring_prefetch(C, R, win, win & 1); cp_async_wait_all(); __syncwarp(); }
    sweep_window_compact<DIR, MODE, WIDE>(C, S, R, X, win, TO);
    __syncwarp();
}

} // namespace age
