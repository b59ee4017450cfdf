// age_kernels.cu -- hand-written sm_100a kernels of the AGE realigner.
//
// Reference semantics (all citations: src/AsmvarAGE/src/AGEaligner.cpp of bioinformatics-centre/AsmVar):
//   K1/K2  calcScores   :978-1083   forward / reverse local fill, single matrix, trace-dependent gap cost
//   K3/K4  calcMaxima   :862-976    running prefix / suffix maxima and their trace
//   K6     findIndelBPs :540-613    K7 findInversionTDuplicationBPs :472-538
//   K8     traceBackMaxima :615-642 K9 addBpoints :644-668
//   K11/12 findLeftBlocks / findRightBlocks :752-860
//
// Layout: one warp owns one *pair* of aligners (two independent DP problems of identical shape in the two
// int16 halves of every register).  A lane holds 8 contig rows in registers, the 32 lanes of the warp form a
// 256-row strip, and the strip is swept along the reference window as an anti-diagonal wavefront: lane p is at
// column t-p in step t and hands its bottom row to lane p+1 with warp shuffles.  Strip boundaries travel
// through a small shared-memory ring that is refilled / flushed with coalesced 512-byte transactions.
#include "age_device.cuh"
#include <cuda_runtime.h>
#include <cstdlib>

namespace age {

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(short)(v & 0xFFFFu); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)(short)(v >> 16); }

template <int DIR> __device__ __forceinline__ uint32_t shfl_prev(uint32_t v)
{
    return DIR > 0 ? __shfl_up_sync(0xFFFFFFFFu, v, 1) : __shfl_down_sync(0xFFFFFFFFu, v, 1);
}

// substitution score (x2) of one half for a special column (Scorer.hh:24-39): codes >= 5 score 0
__device__ __forceinline__ int subst2(int xc, int yc, int m2, int mm2)
{
    if (xc >= 5 || yc >= 5) return 0;
    return xc == yc ? m2 : mm2;
}

// FS code of one cell from its three (2x+tag) candidates -- the tie order of :1015-1017
__device__ __forceinline__ uint32_t fs_code(int d, int h, int v)
{
    int s = d; uint32_t t = T_DIAG;
    if (h >= s) { s = h; t = T_HORI; }
    if (v >= s) { s = v; t = T_VERT; }
    if (s < 0) t = T_STOP;
    return t;
}
// FM code (:899-915 / :952-968): VERTICAL is tested first, HORIZONTAL overrides when strictly greater
__device__ __forceinline__ uint32_t fm_code(int own, int left, int up)
{
    int cur = own; uint32_t t = T_STOP;
    if (up > cur)   { cur = up; t = T_VERT; }
    if (left > cur) { t = T_HORI; }
    return t;
}

// ------------------------------------------------------------------------------------------------
// the sweep: K1+K3 (DIR = +1) or K2+K4 (DIR = -1) of one pair, all strips
// ------------------------------------------------------------------------------------------------
template <int DIR, bool TRACE, bool STOREM>
__device__ void sweep_pair(const PairTask &T, uint4 *ring_in, uint4 *ring_out, const int lane)
{
    constexpr bool FWD = DIR > 0;
    constexpr bool SUM = !FWD;                            // the reverse sweep folds in D = shifted Fmax (K6 :545-550)
    const int p = FWD ? lane : 31 - lane;                 // logical lane: 0 leads the wavefront
    const int len1 = T.len1, len2 = T.len2;
    const uint32_t c0x = T.c0x, kabs = T.kabs, flip = T.flip, gborder = T.gborder;
    const uint32_t *__restrict__ xsel = FWD ? T.xsel_f : T.xsel_r;
    uint2 *__restrict__ Tmat = FWD ? T.Tf : T.Tr;
    uint4 *bnd = FWD ? T.bnd_f : T.bnd_r;
    const int nsteps = len1 + 31;

    for (int si = 0; si < T.nstrips; ++si) {
        const int s = FWD ? si : T.nstrips - 1 - si;
        const int q0 = s * STRIP + lane * RPT;            // 0-based first row of this lane
        const int nvalid = min(max(len2 - q0, 0), RPT);   // rows of this lane that exist
        const size_t sbase = (size_t)s * nsteps;          // first (strip, step) slot of the diagonal-major planes
        uint32_t Ta[RPT], Tb[RPT], Hl[RPT], Gl[RPT], Ml[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const uint2 y = T.ytab[q0 + k];
            Ta[k] = y.x; Tb[k] = y.y;
            Hl[k] = 0; Gl[k] = gborder; Ml[k] = 0;
        }
        uint32_t outH = 0, outG = gborder, outM = 0, outS = 0, prevH = 0;
        const uint4 *bin = si == 0 ? nullptr : bnd + (size_t)(si - 1) * (len1 + 2);
        uint4 *bout = (si + 1 < T.nstrips) ? bnd + (size_t)si * (len1 + 2) : nullptr;
        const bool tail_row = FWD && p == 31 && si + 1 == T.nstrips;   // also stores row q0+8 of D

        auto fetch = [&](int u) -> uint4 {
            if (u > len1) return make_uint4(0, gborder, 0, 0);
            const int c = FWD ? u : len1 + 1 - u;
            if (bin) return __ldcg(bin + c);
            return make_uint4(0, gborder, 0, xsel[c]);
        };
        uint4 pre = fetch(1 + lane);

        // D of the next column, loaded one step ahead (reverse sweep only)
        uint4 Dn0 = make_uint4(0, 0, 0, 0), Dn1 = make_uint4(0, 0, 0, 0);
        auto prefetchD = [&](int un) {
            if (!SUM) return;
            if (un >= 1 && un <= len1) {
                const int c = len1 + 1 - un;
                if (c >= 2) {                              // slot written by forward lane `lane` in step c-1+lane
                    const uint4 *src = reinterpret_cast<const uint4 *>(T.D + (sbase + (size_t)(c - 2 + lane)) * 256 + lane * 8);
                    Dn0 = __ldcg(src); Dn1 = __ldcg(src + 1);
                } else { Dn0 = make_uint4(0, 0, 0, 0); Dn1 = Dn0; }
            }
        };
        prefetchD(1 - p);
        uint32_t best = 0;

        for (int t = 1; t <= nsteps; ++t) {
            if (((t - 1) & 31) == 0) {
                __syncwarp();
                ring_in[lane] = pre;
                __syncwarp();
                pre = fetch(t + 32 + lane);
            }
            uint32_t rH = shfl_prev<DIR>(outH), rG = shfl_prev<DIR>(outG);
            uint32_t rM = shfl_prev<DIR>(outM), rS = shfl_prev<DIR>(outS);
            if (p == 0) { const uint4 e = ring_in[(t - 1) & 31]; rH = e.x; rG = e.y; rM = e.z; rS = e.w; }
            const int u = t - p;
            if (u >= 1 && u <= len1) {
                const int c = FWD ? u : len1 + 1 - u;
                uint32_t Dc[RPT];
                if (SUM) {
                    Dc[0] = Dn0.x; Dc[1] = Dn0.y; Dc[2] = Dn0.z; Dc[3] = Dn0.w;
                    Dc[4] = Dn1.x; Dc[5] = Dn1.y; Dc[6] = Dn1.z; Dc[7] = Dn1.w;
                    prefetchD(u + 1);
                }
                uint32_t S2[RPT];
                if (rS & XSEL_SPECIAL) {
                    const int xa = (rS >> 20) & 7, xb = (rS >> 24) & 7;
                    const int m2a = lo16(T.m2), m2b = hi16(T.m2), mm2a = lo16(T.mm2), mm2b = hi16(T.mm2);
#pragma unroll
                    for (int k = 0; k < RPT; ++k) {
                        const int yc = T.ycode[q0 + k];
                        const int sa = subst2(xa, yc & 15, m2a, mm2a), sb = subst2(xb, yc >> 4, m2b, mm2b);
                        S2[k] = (uint32_t)(sa & 0xFFFF) | ((uint32_t)sb << 16);
                    }
                } else {
                    const uint32_t sel = rS & 0xFFFFu;
#pragma unroll
                    for (int k = 0; k < RPT; ++k) S2[k] = prmt(Ta[k], Tb[k], sel);
                }
                uint32_t vG = rG, vM = rM, dg = prevH;
                uint32_t accA = 0, accB = 0;
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    const int k = FWD ? j : RPT - 1 - j;
                    const uint32_t g  = __vmaxs2(Gl[k], vG);
                    const uint32_t s2 = __viaddmax_s16x2_relu(dg, S2[k], g);      // max(2d, 2h+1, 2v+1, 0)
                    const uint32_t tg = (s2 ^ flip) & 0x00010001u;
                    const uint32_t G2 = __vadd2(tg * kabs + s2, c0x);             // 2*(s + (gap ? ge : go)) + 1
                    const uint32_t Hc = s2 & 0xFFFEFFFEu;
                    const uint32_t Mn = __vimax3_s16x2(Hc, Ml[k], vM);
                    if (TRACE) {
                        const uint32_t d2 = __vadd2(dg, S2[k]);
                        const uint32_t na = fs_code(lo16(d2), lo16(Gl[k]), lo16(vG)) |
                                            (fm_code(lo16(Hc), lo16(Ml[k]), lo16(vM)) << 2);
                        const uint32_t nb = fs_code(hi16(d2), hi16(Gl[k]), hi16(vG)) |
                                            (fm_code(hi16(Hc), hi16(Ml[k]), hi16(vM)) << 2);
                        accA |= na << (4 * k);
                        accB |= nb << (4 * k);
                    }
                    if (SUM) { if (k < nvalid) best = __vmaxu2(best, __vadd2(Mn, Dc[k])); }
                    dg = Hl[k];
                    Hl[k] = Hc; Gl[k] = G2; Ml[k] = Mn;
                    vG = G2; vM = Mn;
                }
                constexpr int last = FWD ? RPT - 1 : 0;
                outH = Hl[last]; outG = Gl[last]; outM = Ml[last]; outS = rS;
                if (STOREM) {
                    if (FWD) {                             // rows q0 .. q0+7 of Fmax (row q0 came from the lane above)
                        uint4 *dst = reinterpret_cast<uint4 *>(T.D + (sbase + (t - 1)) * 256 + lane * 8);
                        dst[0] = make_uint4(rM, Ml[0], Ml[1], Ml[2]);
                        dst[1] = make_uint4(Ml[3], Ml[4], Ml[5], Ml[6]);
                        if (tail_row) T.Dtail[c] = Ml[7];
                    } else {
                        uint4 *dst = reinterpret_cast<uint4 *>(T.Mr + (sbase + (t - 1)) * 256 + lane * 8);
                        dst[0] = make_uint4(Ml[0], Ml[1], Ml[2], Ml[3]);
                        dst[1] = make_uint4(Ml[4], Ml[5], Ml[6], Ml[7]);
                    }
                }
                if (TRACE) Tmat[(sbase + (t - 1)) * 32 + lane] = make_uint2(accA, accB);
                if (p == 31 && bout) ring_out[(u - 1) & 31] = make_uint4(outH, outG, outM, rS);
            }
            prevH = rH;
            if (SUM && ((t & 127) == 0 || t == nsteps)) {
                T.tilemax[((size_t)s * T.nwin + ((t - 1) >> 7)) * 32 + lane] = best;
                best = 0;
            }
            if (bout) {
                const int u31 = t - 31;                   // column just finished by the last logical lane
                if (u31 >= 1 && ((u31 & 31) == 0 || u31 == len1)) {
                    __syncwarp();
                    const int ub = ((u31 - 1) & ~31) + 1 + lane;
                    if (ub <= u31) bout[FWD ? ub : len1 + 1 - ub] = ring_out[lane];
                    __syncwarp();
                }
            }
        }
    }
}

constexpr int SWEEP_WARPS = 4;

template <bool TRACE>
__global__ void __launch_bounds__(SWEEP_WARPS * 32)
age_sweep_kernel(const PairTask *__restrict__ pairs, int n_pairs, int32_t *counter)
{
    __shared__ uint4 rings[SWEEP_WARPS][2][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= n_pairs) break;
        const PairTask &T = pairs[item];
        sweep_pair<+1, TRACE, true>(T, rings[w][0], rings[w][1], lane);
        __threadfence_block();
        __syncwarp();
        sweep_pair<-1, TRACE, TRACE>(T, rings[w][0], rings[w][1], lane);
        __syncwarp();
    }
}

void launch_sweeps(const PairTask *d_pairs, int n_pairs, int32_t *d_counter, cudaStream_t st)
{
    if (n_pairs <= 0) return;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int blocks = (n_pairs + SWEEP_WARPS - 1) / SWEEP_WARPS;
    const int cap = sms * 4;                              // persistent: 4 CTAs x 4 warps per SM
    if (blocks > cap) blocks = cap;
    cudaMemsetAsync(d_counter, 0, sizeof(int32_t), st);
    static const bool fast = getenv("AGE_EXPERIMENT_FAST") != nullptr;     // timing experiment only: no traces
    static const int occ = getenv("AGE_OCC") ? atoi(getenv("AGE_OCC")) : 4;
    if (blocks > sms * occ) blocks = sms * occ;
    if (fast) age_sweep_kernel<false><<<blocks, SWEEP_WARPS * 32, 0, st>>>(d_pairs, n_pairs, d_counter);
    else      age_sweep_kernel<true><<<blocks, SWEEP_WARPS * 32, 0, st>>>(d_pairs, n_pairs, d_counter);
}

// ------------------------------------------------------------------------------------------------
// pass 2: breakpoint enumeration (K5-K9) and traceback (K10-K12) over the materialised matrices
// ------------------------------------------------------------------------------------------------
struct View {
    int len1, len2, nstrips, nsteps, half, flag;
    const uint32_t *D, *Dtail, *Mr;
    const uint2 *Tf, *Tr;

    // The planes are diagonal-major: [strip][step][lane][k].  Forward lane l is at column t-l in step t; reverse
    // lane l (logical lane 31-l) is at column len1+32-t-l.
    __device__ size_t fslot(int q, int c) const { const int l = (q & 255) >> 3; return ((size_t)(q >> 8) * nsteps + (c + l - 1)) * 32 + l; }
    __device__ size_t rslot(int q, int c) const { const int l = (q & 255) >> 3; return ((size_t)(q >> 8) * nsteps + (len1 + 31 - c - l)) * 32 + l; }
    __device__ uint32_t F2w(int r, int c) const {        // packed pair word of 2 * Fmax[r][c]
        if (r < 1 || c < 1 || r > len2 || c > len1) return 0u;
        if ((r >> 8) == nstrips) return Dtail[c];
        return D[fslot(r, c) * 8 + (r & 7)];              // D row index = r (shifted by one row, see the sweep)
    }
    __device__ uint32_t R2w(int r, int c) const {
        if (r < 1 || c < 1 || r > len2 || c > len1) return 0u;
        return Mr[rslot(r - 1, c) * 8 + ((r - 1) & 7)];
    }
    __device__ int F2(int r, int c) const { return (F2w(r, c) >> (16 * half)) & 0xFFFF; }
    __device__ int R2(int r, int c) const { return (R2w(r, c) >> (16 * half)) & 0xFFFF; }
    __device__ uint32_t nibf(int r, int c) const {
        const uint2 w = Tf[fslot(r - 1, c)];
        return ((half ? w.y : w.x) >> (4 * ((r - 1) & 7))) & 15u;
    }
    __device__ uint32_t nibr(int r, int c) const {
        const uint2 w = Tr[rslot(r - 1, c)];
        return ((half ? w.y : w.x) >> (4 * ((r - 1) & 7))) & 15u;
    }
    __device__ bool interior(int r, int c) const { return r >= 1 && c >= 1 && r <= len2 && c <= len1; }
    __device__ uint32_t FS(int r, int c) const { return interior(r, c) ? (nibf(r, c) & 3u) : (uint32_t)T_STOP; }
    __device__ uint32_t RS(int r, int c) const { return interior(r, c) ? (nibr(r, c) & 3u) : (uint32_t)T_STOP; }
    __device__ uint32_t FM(int r, int c) const {         // borders: :874-882
        if (interior(r, c)) return nibf(r, c) >> 2;
        if (r == 0) return c >= 1 ? T_HORI : T_STOP;
        if (c == 0) return r <= len2 ? T_VERT : T_STOP;
        return T_STOP;
    }
    __device__ uint32_t RM(int r, int c) const {         // reverse border keeps RM = 0 (:924-932 sets FM bits there)
        return interior(r, c) ? (nibr(r, c) >> 2) : (uint32_t)T_STOP;
    }
};

struct BpList {
    int n;
    int (*bp)[4];
};

// K8 traceBackMaxima (:615-642)
__device__ void walk_left(const View &V, int &lg1, int &lg2)
{
    uint32_t t;
    while ((t = V.FM(lg2, lg1)) != T_STOP) { if (t == T_VERT) --lg2; else --lg1; }
}
__device__ void walk_right(const View &V, int &rg1, int &rg2)
{
    uint32_t t;
    while ((t = V.RM(rg2, rg1)) != T_STOP) { if (t == T_VERT) ++rg2; else ++rg1; }
}

// K9 addBpoints (:644-668); executed redundantly by every lane on the shared list, lane 0 writes
__device__ int add_bp(const View &V, BpList &L, int lg1, int lg2, int rg1, int rg2, int lane)
{
    if (lg1 == 0 || lg2 == 0) lg1 = lg2 = 0;
    if (rg1 == V.len1 + 1 || lg2 == V.len2 + 1) { rg1 = V.len1 + 1; rg2 = V.len2 + 1; }   // sic: lg2 (:648)
    for (int i = 0; i < L.n; ++i)
        if (L.bp[i][0] - lg1 == L.bp[i][2] - rg1 && L.bp[i][1] - lg2 == L.bp[i][3] - rg2) return 0;
    if (L.n < MAX_BP) {
        if (lane == 0) { int *b = L.bp[L.n]; b[0] = lg1; b[1] = lg2; b[2] = rg1; b[3] = rg2; }
        __syncwarp();
        ++L.n;
        return 1;
    }
    return -1;
}

// K6 findIndelBPs, ordered part (:552-610).  rlo/rhi bound, per row, the columns that can hold a cell with
// Fmax + Rmax == max (from the hot tiles of the reverse sweep and the border rule); rows with rhi < rlo are skipped.
__device__ void indel_enumerate(const View &V, BpList &L, int max2, const int *rlo, const int *rhi, int lane)
{
    const int len2 = V.len2;
    int n_add = 0;
    bool stop = false;
    for (int r = 0; r <= len2 && !stop; ++r) {                           // forward scan (:552-573)
        const int lo = rlo[r], hi = rhi[r];
        if (hi < lo) continue;
        for (int cb = lo & ~31; cb <= hi && !stop; cb += 32) {
            const int c = cb + lane;
            bool hit = false;
            if (c >= lo && c <= hi && V.F2(r, c) + V.R2(r + 1, c + 1) == max2) hit = V.FM(r, c) == T_STOP;
            unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
            while (m && !stop) {
                const int b = __ffs(m) - 1; m &= m - 1;
                int lg1 = cb + b, lg2 = r, rg1 = cb + b + 1, rg2 = r + 1;
                walk_left(V, lg1, lg2); walk_right(V, rg1, rg2);
                const int res = add_bp(V, L, lg1, lg2, rg1, rg2, lane);
                if (res > 0) ++n_add;
                if (n_add >= MAX_BP / 2 || res < 0) stop = true;
            }
        }
    }
    n_add = 0; stop = false;
    for (int r = len2; r >= 0 && !stop; --r) {                            // reverse scan (:586-610)
        const int lo = rlo[r], hi = rhi[r];
        if (hi < lo) continue;
        for (int cb = hi & ~31; cb >= (lo & ~31) && !stop; cb -= 32) {
            const int c = cb + lane;
            bool hit = false;
            if (c >= lo && c <= hi && V.F2(r, c) + V.R2(r + 1, c + 1) == max2) hit = V.RM(r + 1, c + 1) == T_STOP;
            unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
            while (m && !stop) {
                const int b = 31 - __clz(m); m &= ~(1u << b);
                int lg1 = cb + b, lg2 = r, rg1 = cb + b + 1, rg2 = r + 1;
                walk_left(V, lg1, lg2); walk_right(V, rg1, rg2);
                const int res = add_bp(V, L, lg1, lg2, rg1, rg2, lane);
                if (res > 0) ++n_add;
                if (n_add >= MAX_BP / 2 || res < 0) stop = true;
            }
        }
    }
}

// Hot-tile refinement for one half: every reverse-sweep tile whose maximum equals the split maximum is scanned
// exactly and the K6 cells (r, c) = (r'-1, c'-1) with Fmax + Rmax == max widen the column range of their row.
__device__ void indel_row_ranges(const PairTask &T, const View &V, int max2, int *rlo, int *rhi, int lane)
{
    const int len1 = T.len1, len2 = T.len2, h = V.half, nsteps = len1 + 31;
    for (int r = lane; r <= len2; r += 32) { rlo[r] = 0x7FFFFFFF; rhi[r] = -1; }
    __syncwarp();
    const int ntiles = T.nstrips * T.nwin * 32;
    for (int tb = 0; tb < ntiles; tb += 32) {
        const int ti = tb + lane;
        const bool hot = ti < ntiles && (int)((T.tilemax[ti] >> (16 * h)) & 0xFFFF) == max2;
        unsigned m = __ballot_sync(0xFFFFFFFFu, hot);
        while (m) {
            const int tile = tb + __ffs(m) - 1; m &= m - 1;
            const int l = tile & 31, w = (tile >> 5) % T.nwin, s = (tile >> 5) / T.nwin;
            const int q0 = s * STRIP + l * RPT, p = 31 - l;
            for (int i = lane; i < 128; i += 32) {
                const int t = 128 * w + 1 + i;                        // reverse step
                const int u = t - p;                                  // its logical column for lane l
                if (u < 1 || u > len1) continue;
                const int c = len1 + 1 - u;                           // reverse cell column c' (1..len1)
                const uint32_t *mr = T.Mr + (((size_t)s * nsteps + (t - 1)) * 32 + l) * 8;
                const uint32_t *df = T.D + (((size_t)s * nsteps + (c - 2 + l)) * 32 + l) * 8;
                for (int k = 0; k < RPT; ++k) {
                    const int q = q0 + k;                             // reverse cell row r' = q+1
                    if (q >= len2) break;
                    const int rv = (mr[k] >> (16 * h)) & 0xFFFF;
                    const int fv = c >= 2 ? (int)((df[k] >> (16 * h)) & 0xFFFF) : 0;
                    if (rv + fv == max2) { atomicMin(&rlo[q], c - 1); atomicMax(&rhi[q], c - 1); }
                }
            }
        }
    }
    __syncwarp();
    // border cells of K6 (r = len2 or c = len1): Rmax is zero there, so they tie only when Fmax itself == max
    if (V.F2(len2, len1) == max2) {
        // row len2: Fmax[len2][c] is non-decreasing in c -> first column that reaches max
        int lo = 1, hi = len1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (V.F2(len2, mid) >= max2) hi = mid; else lo = mid + 1; }
        if (lane == 0) { rlo[len2] = min(rlo[len2], lo); rhi[len2] = len1; }
        for (int r = 1 + lane; r < len2; r += 32)
            if (V.F2(r, len1) == max2) { rlo[r] = min(rlo[r], len1); rhi[r] = len1; }
    }
    if (max2 == 0) {                                                  // everything ties at zero: scan it all
        for (int r = lane; r <= len2; r += 32) { rlo[r] = 0; rhi[r] = len1; }
    }
    __syncwarp();
}

// K7 findInversionTDuplicationBPs (:472-538).  The reference walks every (j1, j2) pair; here the monotone
// rows are binary-searched for the run of matching values and every H-chain of the maxima trace (whose
// members share one traceBackMaxima endpoint, hence one addBpoints key) is tried once -- the later members
// of a chain are guaranteed duplicates (:652-655) and change nothing.
__device__ void invtdup_enumerate(const View &V, BpList &L, int max2, int lane)
{
    const int len1 = V.len1, len2 = V.len2;
    int n_add = 0;
    bool stop = false;
    for (int r = 0; r <= len2 && !stop; ++r) {                            // forward (:485-504)
        if (V.F2(r, len1) + V.R2(r + 1, 1) != max2) continue;
        for (int jb = 0; jb <= len1 && !stop; jb += 32) {
            const int j = jb + lane;
            unsigned m = __ballot_sync(0xFFFFFFFFu, j <= len1 && V.FM(r, j) == T_STOP);
            while (m && !stop) {
                const int j1 = jb + __ffs(m) - 1; m &= m - 1;
                const int tv = max2 - V.F2(r, j1);
                if (tv < 0) continue;
                // columns x in [1, len1+1] of row r+1, values non-increasing: first x with value <= tv
                int lo = 1, hi = len1 + 2;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (V.R2(r + 1, mid) <= tv) hi = mid; else lo = mid + 1; }
                int x = lo;
                while (x <= len1 + 1 && V.R2(r + 1, x) == tv && !stop) {
                    int xe = x;
                    while (V.RM(r + 1, xe) == T_HORI) ++xe;                // chain shares the endpoint
                    int rg1 = xe, rg2 = r + 1, lg1 = j1, lg2 = r;
                    walk_right(V, rg1, rg2);
                    const int res = add_bp(V, L, lg1, lg2, rg1, rg2, lane);
                    if (res > 0) ++n_add;
                    if (n_add >= MAX_BP / 2 || res < 0) stop = true;
                    x = xe + 1;
                }
            }
        }
    }
    n_add = 0; stop = false;
    for (int i2 = len2 + 1; i2 >= 1 && !stop; --i2) {                     // reverse (:515-534)
        if (V.F2(i2 - 1, len1) + V.R2(i2, 1) != max2) continue;
        for (int jb = ((len1 + 1) / 32) * 32; jb >= 0 && !stop; jb -= 32) {
            const int j = jb + lane;
            unsigned m = __ballot_sync(0xFFFFFFFFu, j >= 1 && j <= len1 + 1 && V.RM(i2, j) == T_STOP);
            while (m && !stop) {
                const int b = 31 - __clz(m); m &= ~(1u << b);
                const int j2 = jb + b;
                const int tv = max2 - V.R2(i2, j2);
                if (tv < 0) continue;
                // columns x in [0, len1] of row i2-1, values non-decreasing: last x with value <= tv
                int lo = -1, hi = len1;
                while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (V.F2(i2 - 1, mid) <= tv) lo = mid; else hi = mid - 1; }
                int x = lo;
                while (x >= 0 && V.F2(i2 - 1, x) == tv && !stop) {
                    int xe = x;
                    while (V.FM(i2 - 1, xe) == T_HORI) --xe;
                    int lg1 = xe, lg2 = i2 - 1, rg1 = j2, rg2 = i2;
                    walk_left(V, lg1, lg2);
                    const int res = add_bp(V, L, lg1, lg2, rg1, rg2, lane);
                    if (res > 0) ++n_add;
                    if (n_add >= MAX_BP / 2 || res < 0) stop = true;
                    x = xe - 1;
                }
            }
        }
    }
}

// K11 findLeftBlocks (:752-808).  dst == nullptr: count only.  Blocks are produced last-first.
__device__ int left_blocks(const View &V, int c, int r, Block *dst, int n_total)
{
    int n = 0, run = 0, last_c = 0, last_r = 0;
    uint32_t t;
    while ((t = V.FS(r, c)) != T_STOP) {
        if (t == T_DIAG) {
            if (run > 0 && last_c - c == 1 && last_r - r == 1) ++run;
            else {
                if (run > 0) { if (dst) dst[n_total - 1 - n] = Block{last_c, last_r, run}; ++n; }
                run = 1;
            }
            last_c = c; last_r = r; --c; --r;
        } else if (t == T_HORI) --c;
        else --r;
    }
    if (run > 0) { if (dst) dst[n_total - 1 - n] = Block{last_c, last_r, run}; ++n; }
    return n;
}

// K12 findRightBlocks (:810-860)
__device__ int right_blocks(const View &V, int c, int r, Block *dst)
{
    int n = 0, run = 0, s1 = 0, s2 = 0;
    uint32_t t;
    while ((t = V.RS(r, c)) != T_STOP) {
        if (t == T_DIAG) {
            if (c <= V.len1 && r <= V.len2) { if (run == 0) { s1 = c; s2 = r; run = 1; } else ++run; }
            ++c; ++r;
        } else {
            if (t == T_HORI) ++c; else ++r;
            if (run > 0) { if (dst) dst[n] = Block{s1, s2, run}; ++n; run = 0; }
        }
    }
    if (run > 0) { if (dst) dst[n] = Block{s1, s2, run}; ++n; }
    return n;
}

constexpr int P2_WARPS = 4;

__global__ void __launch_bounds__(P2_WARPS * 32)
age_pass2_kernel(Pass2Params prm)
{
    __shared__ int bp_s[P2_WARPS][MAX_BP][4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(prm.next, 1);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= prm.n_pairs) break;
        const PairTask &T = prm.pairs[item];
        const int len1 = T.len1, len2 = T.len2;

        // ---- split maxima of both halves (:545-550 / :476-481) ------------------------------------------
        uint32_t mx_indel = 0;                                   // packed u16x2: max of 2*(Fmax + Rmax)
        const int ntiles = T.nstrips * T.nwin * 32;
        for (int i = lane; i < ntiles; i += 32) mx_indel = __vmaxu2(mx_indel, T.tilemax[i]);
        const View V0{len1, len2, T.nstrips, len1 + 31, 0, 0, T.D, T.Dtail, T.Mr, T.Tf, T.Tr};
        uint32_t mx_col = 0;
        if (T.flag[0] != F_INDEL || T.flag[1] != F_INDEL)
            for (int r = lane; r <= len2; r += 32) mx_col = __vmaxu2(mx_col, __vadd2(V0.F2w(r, len1), V0.R2w(r + 1, 1)));
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            mx_indel = __vmaxu2(mx_indel, __shfl_xor_sync(0xFFFFFFFFu, mx_indel, o));
            mx_col   = __vmaxu2(mx_col,   __shfl_xor_sync(0xFFFFFFFFu, mx_col, o));
        }
        mx_indel = __vmaxu2(mx_indel, V0.F2w(len2, len1));                          // Fmax[len2][len1] + 0 (border cells)
        int max2[2];
        for (int h = 0; h < 2; ++h) {
            const uint32_t v = (T.flag[h] == F_INDEL) ? mx_indel : mx_col;
            max2[h] = (v >> (16 * h)) & 0xFFFF;
        }
        bool need[2] = {T.out[0] >= 0 && T.flag[0] != 0, T.out[1] >= 0 && T.flag[1] != 0};
        if (T.select && need[0] && need[1]) {                    // age_align.cpp:273 / AGEaligner.cpp:151: ties keep half 0
            const int win = max2[1] > max2[0] ? 1 : 0;
            if (lane == 0) {
                AlignerOut &Lo = prm.outs[T.out[1 - win]];
                Lo.status = 0; Lo.score = max2[1 - win] >> 1; Lo.n_bp = -1; Lo.alt_index = -1;
                for (int q = 0; q < 4; ++q) { Lo.n_blk[q] = 0; Lo.blk_off[q] = 0; }
            }
            need[1 - win] = false;
        }

        for (int h = 0; h < 2; ++h) {
            if (!need[h]) continue;
            View V{len1, len2, T.nstrips, len1 + 31, h, T.flag[h], T.D, T.Dtail, T.Mr, T.Tf, T.Tr};
            AlignerOut &O = prm.outs[T.out[h]];
            BpList L{0, bp_s[w]};
            __syncwarp();
            if (T.flag[h] == F_INDEL) {
                int *rlo = T.rowrange, *rhi = T.rowrange + (len2 + 2);
                indel_row_ranges(T, V, max2[h], rlo, rhi, lane);
                indel_enumerate(V, L, max2[h], rlo, rhi, lane);
            } else invtdup_enumerate(V, L, max2[h], lane);
            __syncwarp();
            // ---- K10: blocks of bpoint 0 and of the alternative shown by printAlignment (:287-306) ----
            if (lane == 0) {
                O.status = 0;
                O.score = max2[h] >> 1;
                O.n_bp = L.n;
                for (int i = 0; i < L.n; ++i)
                    for (int k = 0; k < 4; ++k) O.bp[i][k] = L.bp[i][k];
                int alt = -1;
                for (int which = 0; which < 2; ++which) {
                    int idx = 0;
                    int nl = 0, nr = 0;
                    if (which == 0) {
                        nl = left_blocks(V, L.bp[0][0], L.bp[0][1], nullptr, 0);
                        nr = right_blocks(V, L.bp[0][2], L.bp[0][3], nullptr);
                    } else {
                        for (idx = L.n - 1; idx >= 1; --idx) {
                            nl = left_blocks(V, L.bp[idx][0], L.bp[idx][1], nullptr, 0);
                            nr = right_blocks(V, L.bp[idx][2], L.bp[idx][3], nullptr);
                            if (nl + nr > 0) break;
                        }
                        if (idx < 1) { nl = nr = 0; idx = -1; }
                        alt = idx;
                    }
                    int off = 0;
                    if (nl + nr > 0) {
                        off = atomicAdd(prm.pool_used, nl + nr);
                        if (off + nl + nr <= prm.pool_cap) {
                            left_blocks(V, L.bp[idx][0], L.bp[idx][1], prm.pool + off, nl);
                            right_blocks(V, L.bp[idx][2], L.bp[idx][3], prm.pool + off + nl);
                        } else O.status = -4;                   // pool overflow: host retries with a larger pool
                    }
                    O.n_blk[2 * which] = nl; O.n_blk[2 * which + 1] = nr;
                    O.blk_off[2 * which] = off; O.blk_off[2 * which + 1] = off + nl;
                }
                O.alt_index = alt;
            }
            __syncwarp();
        }
    }
}

void launch_pass2(const Pass2Params &p, cudaStream_t st)
{
    if (p.n_pairs <= 0) return;
    if (getenv("AGE_EXPERIMENT_FAST")) return;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int blocks = (p.n_pairs + P2_WARPS - 1) / P2_WARPS;
    const int cap = sms * 4;
    if (blocks > cap) blocks = cap;
    cudaMemsetAsync(p.next, 0, sizeof(int32_t), st);
    age_pass2_kernel<<<blocks, P2_WARPS * 32, 0, st>>>(p);
}

} // namespace age
