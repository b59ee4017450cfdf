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
    const int p = DIR > 0 ? lane : 31 - lane;            // logical lane: 0 leads the wavefront
    const int len1 = T.len1;
    const int P = T.P;
    const uint32_t c0x = T.c0x, kabs = T.kabs, flip = T.flip, gborder = T.gborder;
    const uint32_t *__restrict__ xsel = DIR > 0 ? T.xsel_f : T.xsel_r;
    uint32_t *__restrict__ Mmat = DIR > 0 ? T.Mf : T.Mr;
    uint2 *__restrict__ Tmat = DIR > 0 ? T.Tf : T.Tr;
    uint4 *bnd = DIR > 0 ? T.bnd_f : T.bnd_r;
    const int nsteps = len1 + 31;

    for (int si = 0; si < T.nstrips; ++si) {
        const int s = DIR > 0 ? si : T.nstrips - 1 - si;
        const int q0 = s * STRIP + lane * RPT;           // 0-based first row of this lane
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

        auto fetch = [&](int u) -> uint4 {
            if (u > len1) return make_uint4(0, gborder, 0, 0);
            const int c = DIR > 0 ? u : len1 + 1 - u;
            if (bin) return bin[c];
            return make_uint4(0, gborder, 0, xsel[c]);
        };
        uint4 pre = fetch(1 + lane);

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
                const int c = DIR > 0 ? u : len1 + 1 - u;
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
                    const int k = DIR > 0 ? j : RPT - 1 - j;
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
                    dg = Hl[k];
                    Hl[k] = Hc; Gl[k] = G2; Ml[k] = Mn;
                    vG = G2; vM = Mn;
                }
                constexpr int last = DIR > 0 ? RPT - 1 : 0;
                outH = Hl[last]; outG = Gl[last]; outM = Ml[last]; outS = rS;
                if (STOREM) {
                    uint4 *dst = reinterpret_cast<uint4 *>(Mmat + (size_t)c * P + q0);
                    dst[0] = make_uint4(Ml[0], Ml[1], Ml[2], Ml[3]);
                    dst[1] = make_uint4(Ml[4], Ml[5], Ml[6], Ml[7]);
                }
                if (TRACE) Tmat[(size_t)c * (P / RPT) + (q0 / RPT)] = make_uint2(accA, accB);
                if (p == 31 && bout) ring_out[(u - 1) & 31] = make_uint4(outH, outG, outM, rS);
            }
            prevH = rH;
            if (bout) {
                const int u31 = t - 31;                   // column just finished by the last logical lane
                if (u31 >= 1 && ((u31 & 31) == 0 || u31 == len1)) {
                    __syncwarp();
                    const int ub = ((u31 - 1) & ~31) + 1 + lane;
                    if (ub <= u31) bout[DIR > 0 ? ub : len1 + 1 - ub] = ring_out[lane];
                    __syncwarp();
                }
            }
        }
    }
}

constexpr int SWEEP_WARPS = 4;

__global__ void __launch_bounds__(SWEEP_WARPS * 32)
age_sweep_kernel(const PairTask *__restrict__ pairs, int n_pairs, int32_t *counter)
{
    __shared__ uint4 rings[SWEEP_WARPS][2][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= 2 * n_pairs) break;
        const PairTask &T = pairs[item >> 1];
        if (item & 1) sweep_pair<-1, true, true>(T, rings[w][0], rings[w][1], lane);
        else          sweep_pair<+1, true, true>(T, rings[w][0], rings[w][1], lane);
        __syncwarp();
    }
}

void launch_sweeps(const PairTask *d_pairs, int n_pairs, int32_t *d_counter, cudaStream_t st)
{
    if (n_pairs <= 0) return;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int items = 2 * n_pairs;
    int blocks = (items + SWEEP_WARPS - 1) / SWEEP_WARPS;
    const int cap = sms * 4;                              // persistent: 4 CTAs x 4 warps per SM
    if (blocks > cap) blocks = cap;
    cudaMemsetAsync(d_counter, 0, sizeof(int32_t), st);
    age_sweep_kernel<<<blocks, SWEEP_WARPS * 32, 0, st>>>(d_pairs, n_pairs, d_counter);
}

// ------------------------------------------------------------------------------------------------
// pass 2: breakpoint enumeration (K5-K9) and traceback (K10-K12) over the materialised matrices
// ------------------------------------------------------------------------------------------------
struct View {
    int len1, len2, P, half, flag;
    const uint32_t *Mf, *Mr;
    const uint2 *Tf, *Tr;

    __device__ int F2(int r, int c) const {              // 2 * Fmax[r][c], zero on the borders
        if (r < 1 || c < 1 || r > len2 || c > len1) return 0;
        return (Mf[(size_t)c * P + (r - 1)] >> (16 * half)) & 0xFFFF;
    }
    __device__ int R2(int r, int c) const {              // 2 * Rmax[r][c]
        if (r < 1 || c < 1 || r > len2 || c > len1) return 0;
        return (Mr[(size_t)c * P + (r - 1)] >> (16 * half)) & 0xFFFF;
    }
    __device__ uint32_t nib(const uint2 *T, int r, int c) const {
        const uint2 w = T[(size_t)c * (P / RPT) + ((r - 1) >> 3)];
        return ((half ? w.y : w.x) >> (4 * ((r - 1) & 7))) & 15u;
    }
    __device__ bool interior(int r, int c) const { return r >= 1 && c >= 1 && r <= len2 && c <= len1; }
    __device__ uint32_t FS(int r, int c) const { return interior(r, c) ? (nib(Tf, r, c) & 3u) : (uint32_t)T_STOP; }
    __device__ uint32_t RS(int r, int c) const { return interior(r, c) ? (nib(Tr, r, c) & 3u) : (uint32_t)T_STOP; }
    __device__ uint32_t FM(int r, int c) const {         // borders: :874-882
        if (interior(r, c)) return nib(Tf, r, c) >> 2;
        if (r == 0) return c >= 1 ? T_HORI : T_STOP;
        if (c == 0) return r <= len2 ? T_VERT : T_STOP;
        return T_STOP;
    }
    __device__ uint32_t RM(int r, int c) const {         // reverse border keeps RM = 0 (:924-932 sets FM bits there)
        return interior(r, c) ? (nib(Tr, r, c) >> 2) : (uint32_t)T_STOP;
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

// K6 findIndelBPs, ordered part (:552-610) restricted to the rows flagged by the coalesced pre-scan
__device__ void indel_enumerate(const View &V, BpList &L, int max2, const uint32_t *rowmask, int lane)
{
    const int len1 = V.len1, len2 = V.len2;
    const uint32_t fbit = 1u << (2 * V.half), rbit = 2u << (2 * V.half);
    int n_add = 0;
    bool stop = false;
    for (int r = 0; r <= len2 && !stop; ++r) {                           // forward scan (:552-573)
        if (!(rowmask[r] & fbit)) continue;
        for (int cb = 0; cb <= len1 && !stop; cb += 32) {
            const int c = cb + lane;
            bool hit = false;
            if (c <= len1 && V.F2(r, c) + V.R2(r + 1, c + 1) == max2) hit = V.FM(r, c) == T_STOP;
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
        if (!(rowmask[r] & rbit)) continue;
        for (int cb = (len1 / 32) * 32; cb >= 0 && !stop; cb -= 32) {
            const int c = cb + lane;
            bool hit = false;
            if (c <= len1 && V.F2(r, c) + V.R2(r + 1, c + 1) == max2) hit = V.RM(r + 1, c + 1) == T_STOP;
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
        const int len1 = T.len1, len2 = T.len2, P = T.P;

        // ---- coalesced pre-scan: split maxima of both halves (:545-550 / :476-481) -----------------
        uint32_t mx_indel = 0;                                   // packed u16x2: max of 2*(Fmax + Rmax)
        for (int rb = 0; rb <= len2; rb += 32) {
            const int r = rb + lane;
            if (r > len2) continue;
            for (int c = 0; c <= len1; ++c) {
                const uint32_t f = (r >= 1 && c >= 1) ? T.Mf[(size_t)c * P + (r - 1)] : 0u;
                const uint32_t q = (r + 1 <= len2 && c + 1 <= len1) ? T.Mr[(size_t)(c + 1) * P + r] : 0u;
                mx_indel = __vmaxu2(mx_indel, __vadd2(f, q));
            }
        }
        uint32_t mx_col = 0;
        for (int r = lane; r <= len2; r += 32) {
            const uint32_t f = r >= 1 ? T.Mf[(size_t)len1 * P + (r - 1)] : 0u;
            const uint32_t q = r + 1 <= len2 ? T.Mr[(size_t)1 * P + r] : 0u;
            mx_col = __vmaxu2(mx_col, __vadd2(f, q));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            mx_indel = __vmaxu2(mx_indel, __shfl_xor_sync(0xFFFFFFFFu, mx_indel, o));
            mx_col   = __vmaxu2(mx_col,   __shfl_xor_sync(0xFFFFFFFFu, mx_col, o));
        }
        int max2[2];
        for (int h = 0; h < 2; ++h) {
            const uint32_t v = (T.flag[h] == F_INDEL) ? mx_indel : mx_col;
            max2[h] = (v >> (16 * h)) & 0xFFFF;
        }
        // ---- coalesced pre-scan 2: rows that hold a cell with Fmax + Rmax == max (indel halves) -------
        const bool indelA = T.flag[0] == F_INDEL && T.out[0] >= 0, indelB = T.flag[1] == F_INDEL && T.out[1] >= 0;
        if (indelA || indelB) {
            const uint32_t want = (uint32_t)max2[0] | ((uint32_t)max2[1] << 16);
            for (int rb = 0; rb <= len2; rb += 32) {
                const int r = rb + lane;
                if (r > len2) continue;
                uint32_t bits = 0;
                for (int c = 0; c <= len1; ++c) {
                    const uint32_t f = (r >= 1 && c >= 1) ? T.Mf[(size_t)c * P + (r - 1)] : 0u;
                    const uint32_t q = (r + 1 <= len2 && c + 1 <= len1) ? T.Mr[(size_t)(c + 1) * P + r] : 0u;
                    const uint32_t e = __vadd2(f, q) ^ want;
                    if ((e & 0xFFFFu) == 0) bits |= 3u;
                    if ((e >> 16) == 0) bits |= 12u;
                }
                T.rowmask[r] = bits;
            }
            __syncwarp();
        }

        for (int h = 0; h < 2; ++h) {
            if (T.out[h] < 0 || T.flag[h] == 0) continue;
            View V{len1, len2, P, h, T.flag[h], T.Mf, T.Mr, T.Tf, T.Tr};
            AlignerOut &O = prm.outs[T.out[h]];
            BpList L{0, bp_s[w]};
            __syncwarp();
            if (T.flag[h] == F_INDEL) indel_enumerate(V, L, max2[h], T.rowmask, lane);
            else                      invtdup_enumerate(V, L, max2[h], lane);
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
