// age_pass2.cuh -- pass 2 of a pair: split maximum and partner selection, breakpoint enumeration (K5-K9) and the
// tracebacks (K10-K12), as warp-collective device functions.  One warp works on one aligner; every trace bit it looks at
// comes from re-running the sweep over the one window that holds it (sweep_window_trace), from that window's checkpoint.
//
// Reference semantics (src/AsmvarAGE/src/AGEaligner.cpp): K6 findIndelBPs :540-613, K7 findInversionTDuplicationBPs
// :472-538, K8 traceBackMaxima :615-642, K9 addBpoints :644-668, K11/K12 findLeftBlocks / findRightBlocks :752-860.
#pragma once
#include "age_sweep.cuh"

// AGE_K7_LADDER: how many VERTICAL steps merges_into_next follows before it gives up and lets the walk be made (any value is exact)
#ifndef AGE_K7_LADDER
#define AGE_K7_LADDER 64
#endif
// AGE_V_WIDERUN (build switch for A/B measurements): 1 = straight runs of the maxima planes are read 256 cells per pass, 0 = 32
#ifndef AGE_V_WIDERUN
#define AGE_V_WIDERUN 1
#endif

namespace age {

// optional cycle accounting of pass 2 (compiled in with -DAGE_P2_PROF only): 1 ranges / K7 forward, 2 enumerate / K7 reverse,
// 3 blocks, 4 forward re-sweeps, 5 reverse re-sweeps
#ifdef AGE_P2_PROF
__device__ unsigned long long g_prof[16];      // 8 + plane: on-demand re-sweeps per plane (FS, FM, RS, RM), 12 + plane: with the maxima (K7)
constexpr int PROF_PAIRS = 1 << 15;
// per pair of the launch: cycles in the enumeration, in the queueing of the traceback paths, in the blocks kernel; windows re-swept
// on demand by the enumerate / by the blocks kernel
__device__ unsigned int g_pair_cycles[12][PROF_PAIRS];    // + 5: cycles in the row ranges, 6: hot windows, 7: cycles in the forward scan (K6)
#define P2_PROF_PAIR(i, item, v) do { if ((threadIdx.x & 31) == 0 && (item) < PROF_PAIRS) g_pair_cycles[i][item] += (unsigned int)(clock64() - (v)); } while (0)
__device__ unsigned long long g_prof2[8];      // K7: 0 tie rows, 1 / 2 forward / reverse start cells, 3 / 4 cycles spent in them, 5 cycles in ensure_rm
#define P2_PROF2_ADD(i, val) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_prof2[i], (unsigned long long)(val)); } while (0)
#define P2_PROF_T0(v) long long v = clock64()
#define P2_OD(V, i, val) do { if ((threadIdx.x & 31) == 0 && (V).od) (V).od[i] += (unsigned int)(val); } while (0)
#define P2_PROF_RESET(v) v = clock64()
#define P2_PROF_ADD(i, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_prof[i], (unsigned long long)(clock64() - (v))); } while (0)
#else
#define P2_PROF_T0(v) do { } while (0)
#define P2_PROF_RESET(v) do { } while (0)
#define P2_PROF_ADD(i, v) do { } while (0)
#define P2_PROF_PAIR(i, item, v) do { } while (0)
#define P2_PROF2_ADD(i, val) do { } while (0)
#define P2_OD(V, i, val) do { } while (0)
#endif

template <bool WIDE> struct View {
    const PairTask *T;
    P2Scratch sc;
    WarpRings *R;
    TraceSmem *X;
    int len1, len2, nb, nsteps, nwin, nstrips, half, lane, max2, vwords;
#ifdef AGE_P2_PROF
    unsigned int *od = nullptr;                  // [4] of the pair at hand: windows re-swept on demand, cycles in the row ranges, hot windows, cycles in the forward scan
#endif
    __device__ __forceinline__ int sh() const { return WIDE ? 0 : 16 * half; }                                  // the aligner's half of a packed word
    __device__ __forceinline__ int pick(uint32_t w) const { return WIDE ? (int)w : (int)((w >> (16 * half)) & 0xFFFFu); }

    // ---- geometry --------------------------------------------------------------------------------
    __device__ int fstep(int q, int c) const { return (c - 1) / W + 1 + ((q & 255) >> 3); }          // forward step of cell (q+1, c)
    __device__ int rstep(int q, int c) const { return nb - (c - 1) / W + 31 - ((q & 255) >> 3); }    // reverse step
    __device__ bool interior(int r, int c) const { return r >= 1 && c >= 1 && r <= len2 && c <= len1; }

    // ---- pages of the trace pool (P2Scratch) -------------------------------------------------------------
    // page of `plane` (0 .. 3 codes, 4 = Mp) that holds step t of the strip of row q (no residency check)
    __device__ __forceinline__ char *page(int plane, int q, int t) const {
        return sc.pool + (size_t)sc.ptab[plane * sc.nwt + (q >> 8) * nwin + ((t - 1) >> 5)] * P2_GRANULE;
    }
    __device__ __forceinline__ int in_page(int q, int t) const { return ((t - 1) & 31) * 32 + ((q & 255) >> 3); }   // [step in window][lane]
    __device__ __forceinline__ uint2 code_word(int plane, int q, int t) const { return reinterpret_cast<const uint2 *>(page(plane, q, t))[in_page(q, t)]; }
    __device__ __forceinline__ uint32_t hot_at(int q, int t) const { return reinterpret_cast<const uint32_t *>(page(3, q, t) + 2 * P2_GRANULE)[in_page(q, t)]; }
    // take a page for (plane, strip s, window win) from the pool: lane 0 bumps the counter; once the pool is exhausted every
    // further page is the dump at its start and the host runs the batch again with a larger pool
    __device__ char *take_page(int plane, int s, int win) const {
        uint32_t g = 0;
        if (lane == 0) {
            const uint32_t n = (uint32_t)p2_page_granules(plane);
            g = atomicAdd(sc.pool_next, n) + P2_DUMP_GRANULES;
            if (g + n > sc.pool_granules) g = 0;
            sc.ptab[plane * sc.nwt + s * nwin + win] = (int32_t)g;
        }
        g = __shfl_sync(0xFFFFFFFFu, g, 0);
        return sc.pool + (size_t)g * P2_GRANULE;
    }
    __device__ TraceOut trace_out(int plane, int s, int win, bool with_mp) const {
        char *pg = take_page(plane, s, win);
        TraceOut TO{reinterpret_cast<uint2 *>(pg), reinterpret_cast<uint32_t *>(pg + 2 * P2_GRANULE), nullptr, sh(), max2};
        if (with_mp) TO.Mp = reinterpret_cast<uint32_t *>(take_page(4, s, win));
        return TO;
    }

    // ---- on-demand trace windows ---------------------------------------------------------------------
    // planes: 0 = FS, 1 = FM, 2 = RS, 3 = RM (+ hot bits), 4 = reverse maxima (K7)
    __device__ bool valid(int plane, int s, int win) const {
        const int i = s * nwin + win;
        return (sc.valid[plane * vwords + (i >> 5)] >> (i & 31)) & 1u;
    }
    __device__ void set_valid(int plane, int s, int win) const {
        const int i = s * nwin + win;
        __syncwarp();
        if (lane == 0) sc.valid[plane * vwords + (i >> 5)] |= 1u << (i & 31);
        __syncwarp();
    }
    // (not inlined: one copy of the four compact re-sweeps per kernel, with the lane state in registers)
    __device__ __noinline__ void ensure(int plane, int s, int win, bool with_mp = false) const {
        if (valid(with_mp ? 4 : plane, s, win)) return;
        P2_PROF_T0(c0);
        const TraceOut TO = trace_out(plane, s, win, with_mp);
        switch (plane) {
        case 0:  sweep_window_trace<+1, MODE_TRACE_S, WIDE>(*T, *R, *X, lane, s, win, TO); break;
        case 1:  sweep_window_trace<+1, MODE_TRACE_M, WIDE>(*T, *R, *X, lane, s, win, TO); break;
        case 2:  sweep_window_trace<-1, MODE_TRACE_S, WIDE>(*T, *R, *X, lane, s, win, TO); break;
        default: sweep_window_trace<-1, MODE_TRACE_M, WIDE>(*T, *R, *X, lane, s, win, TO); break;
        }
        set_valid(plane, s, win);
        if (with_mp) set_valid(4, s, win);
        P2_PROF_ADD(4 + (plane >> 1), c0);
#ifdef AGE_P2_PROF
        if (lane == 0) { atomicAdd(&g_prof[6 + (plane >> 1)], 1ull); atomicAdd(&g_prof[(with_mp ? 12 : 8) + plane], 1ull); if (od) ++*od; }
#endif
    }
    __device__ void ensure_rm(int s) const {                 // reverse maxima of a whole strip (K7 rows)
        for (int win = 0; win < nwin; ++win) ensure(3, s, win, true);
    }
    __device__ void reset_valid() const {
        __syncwarp();
        for (int i = lane; i < 5 * vwords; i += 32) sc.valid[i] = 0;
        __syncwarp();
    }
    __device__ int win_of(int plane, int r, int c) const { return (((plane >> 1) ? rstep(r - 1, c) : fstep(r - 1, c)) - 1) / CKS; }
    // Every lane names one cell of `plane`; windows that are not resident are re-swept one after the other until
    // every lane's cell can be read (warp-collective).
    __device__ void ensure_lanes(int plane, int r, int c) const {
        const bool in = interior(r, c);
        const int s = in ? ((r - 1) >> 8) : 0, win = in ? win_of(plane, r, c) : 0;
        bool ok = !in || valid(plane, s, win);
        unsigned m;
        while ((m = __ballot_sync(0xFFFFFFFFu, !ok)) != 0) {
            const int src = __ffs(m) - 1;
            const int ss = __shfl_sync(0xFFFFFFFFu, s, src), ww = __shfl_sync(0xFFFFFFFFu, win, src);
            ensure(plane, ss, ww);
            ok = ok || (s == ss && win == ww);
        }
    }

    // ---- maxima ------------------------------------------------------------------------------------------
    __device__ uint32_t F2w(int r, int c) const {        // packed pair word of 2 * Fmax[r][c]
        if (r < 1 || c < 1 || r > len2 || c > len1) return 0u;
        if (c == len1) return T->Dlast[r];
        if ((r >> 8) == nstrips) return T->Dtail[c - 1];
        const int l = (r & 255) >> 3, k = r & 7, t = c / W + 1 + l, w = c % W;   // D is shifted by one row and column
        const size_t slot = (size_t)(r >> 8) * nsteps + (t - 1);
        if (WIDE || T->dfmt == DFMT_RAW) return T->D[(((slot * W + w) * 2 + (k >> 2)) * 32 + l) * 4 + (k & 3)];
        const uint32_t *q = T->D + (slot * 5 * 32 + l) * 4;                      // OFF8: entry 0 of the column + the row's offset bytes
        const uint32_t wd = q[(size_t)(1 + w) * 128 + (k >> 1)] >> (8 * (k & 1));       // bytes {A_2j, A_2j+1, B_2j, B_2j+1}
        return q[w] + (wd & 0x00FF00FFu);
    }
    __device__ int F2(int r, int c) const { return pick(F2w(r, c)); }
    __device__ int R2(int r, int c) const {              // 2 * Rmax[r][c]; needs ensure_rm(strip of r)
        if (r < 1 || c < 1 || r > len2 || c > len1) return 0;
        if (c == 1) return pick(T->Rfirst[r]);
        const int q = r - 1, l = (q & 255) >> 3, k = q & 7, t = rstep(q, c), w = W * ((c - 1) / W + 1) - c;
        return pick(reinterpret_cast<const uint32_t *>(page(4, q, t))[((((((t - 1) & 31) * W + w) * 2 + (k >> 2)) * 32 + l) * 4) + (k & 3)]);
    }

    // ---- trace codes (no residency check: callers ensure first) ---------------------------------------------------
    __device__ uint32_t FM_border(int r, int c) const {  // :874-882
        if (r == 0) return c >= 1 ? T_HORI : T_STOP;
        if (c == 0) return r <= len2 ? T_VERT : T_STOP;
        return T_STOP;
    }
    // code of plane 0 = FS, 1 = FM, 2 = RS, 3 = RM at any cell (cells outside the matrix stop a chain; the reverse
    // border keeps RM = 0, :924-932)
    __device__ uint32_t code_nc(int plane, int r, int c) const {
        if (r < 0 || c < 0 || r > len2 + 1 || c > len1 + 1) return T_STOP;
        if (!interior(r, c)) return plane == 1 ? FM_border(r, c) : (uint32_t)T_STOP;
        const int q = r - 1, rev = plane >> 1;
        const uint2 v = code_word(plane, q, rev ? rstep(q, c) : fstep(q, c));
        const int w = rev ? W * ((c - 1) / W + 1) - c : (c - 1) % W;
        return (((w < 2 ? v.x : v.y) >> (16 * (w & 1))) >> (2 * (q & 7))) & 3u;
    }
    __device__ bool hot_nc(int r, int c) const {          // reverse cell (r, c): Fmax[r-1][c-1] + Rmax[r][c] == max
        const int q = r - 1;
        const int w = W * ((c - 1) / W + 1) - c;
        return (hot_at(q, rstep(q, c)) >> (8 * w + (q & 7))) & 1u;
    }
    // Row scans, one block of four columns per lane (128 cells per pass instead of 32).  A plane word holds the codes of
    // 8 rows x 4 columns of one (lane, step) slot; word_codes() picks the row's four codes, 2 bits per block column w in
    // processing order (forward planes: w = column offset in the block; reverse planes: w = 3 - offset).
    __device__ __forceinline__ uint32_t word_codes(uint2 v, int q) const {
        const int sh = 2 * (q & 7);
        return ((v.x >> sh) & 3u) | (((v.x >> (16 + sh)) & 3u) << 2) | (((v.y >> sh) & 3u) << 4) | (((v.y >> (16 + sh)) & 3u) << 6);
    }
    // codes of FM (plane 1) or FS (plane 0) at row r (1-based, interior) of the forward block with columns c1 .. c1+3 (c1 = 4m + 1)
    __device__ __forceinline__ uint32_t fwd_codes(int plane, int r, int c1) const { return word_codes(code_word(plane, r - 1, fstep(r - 1, c1)), r - 1); }
    // codes of RM (plane 3) or RS (plane 2) at row r of the reverse block with columns c1 .. c1+3 (c1 = 4m + 1): column c1 + j is w = 3 - j
    __device__ __forceinline__ uint32_t rev_codes(int plane, int r, int c1) const { return word_codes(code_word(plane, r - 1, rstep(r - 1, c1)), r - 1); }
    __device__ __forceinline__ uint32_t hot_word(int r, int c) const { return hot_at(r - 1, rstep(r - 1, c)); }     // of reverse cell (r, c)'s slot

    // Starting at (r, c) and stepping by (dr, dc): how many consecutive cells carry code t, and the code of the first cell
    // that does not -- or RUN_ON when the look-ahead ended without finding one (the caller reads on from the cell behind it).
    // Only the cell the walk stands on forces its window to be swept; cells further along are read when their window is
    // resident and end the look-ahead otherwise: a walk that stops a few cells short of a window border must not pay a
    // whole re-sweep (hundreds of thousands of cycles inside this latency-bound kernel) for cells it never reaches.
    // Straight runs of the maxima planes (FM / RM: a walk crosses the whole excised stretch along one row or column) read
    // one trace word per lane and word instead of one cell: along a row two blocks of four columns per lane = 256 cells a
    // pass, along a column the eight rows a word holds = 256 cells; everything else, and anything near the border, 32 cells.
    static constexpr uint32_t RUN_ON = 0xFFu;
    __device__ __forceinline__ bool resident(int plane, int r, int c) const { return valid(plane, (r - 1) >> 8, win_of(plane, r, c)); }   // interior (r, c)
    __device__ int run_len(int plane, int r, int c, int dr, int dc, uint32_t t, uint32_t &tnext) const {
        const int rev = plane >> 1;
        if (interior(r, c)) ensure(plane, (r - 1) >> 8, win_of(plane, r, c));
#if AGE_V_WIDERUN
        if ((plane & 1) && dr == 0 && r >= 1 && r <= len2 && c >= 1 && c <= len1) {      // along row r: blocks m0, m0 + dc, ... (columns 4m+1 .. 4m+4)
            const int m0 = (c - 1) >> 2, mlast = m0 + dc * 63;
            if (mlast >= 0 && 4 * mlast + 4 <= len1) {
                const int q = r - 1;
                const int first = dc > 0 ? 4 - ((c - 1) & 3) : ((c - 1) & 3) + 1;       // cells of the run in block m0
                int brk = -1; uint32_t bcode = RUN_ON;                                // first cell of this lane that ends the look-ahead
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int i = 2 * lane + g, m = m0 + dc * i, c1 = 4 * m + 1;
                    const int before = i == 0 ? 0 : first + 4 * (i - 1), cnt = i == 0 ? first : 4;
                    if (brk >= 0) continue;
                    if (!resident(plane, r, c1)) { brk = before; continue; }           // (bcode stays RUN_ON)
                    const uint32_t codes = word_codes(code_word(plane, q, rev ? rstep(q, c1) : fstep(q, c1)), q);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // k-th cell of the block in walking order: column offset j; reverse planes count block columns from the right
                        const int j = dc > 0 ? (i == 0 ? 4 - first + k : k) : (i == 0 ? first - 1 - k : 3 - k);
                        const uint32_t code = (codes >> (2 * (rev ? 3 - j : j))) & 3u;
                        if (k < cnt && brk < 0 && code != t) { brk = before + k; bcode = code; }
                    }
                }
                const unsigned mk = __ballot_sync(0xFFFFFFFFu, brk >= 0);
                if (!mk) { tnext = RUN_ON; return first + 4 * 63; }
                const int src = __ffs(mk) - 1;
                tnext = __shfl_sync(0xFFFFFFFFu, bcode, src);
                return __shfl_sync(0xFFFFFFFFu, brk, src);
            }
        }
        if ((plane & 1) && dc == 0 && c >= 1 && c <= len1 && r >= 1 && r <= len2) {      // along column c: groups of the eight rows one word holds
            const int q0 = r - 1, g0 = q0 >> 3, glast = g0 + dr * 31;
            if (glast >= 0 && 8 * glast + 8 <= len2) {
                const int first = dr > 0 ? 8 - (q0 & 7) : (q0 & 7) + 1;
                const int qg = 8 * (g0 + dr * lane);                                  // first row (0-based) of this lane's group
                const int before = lane == 0 ? 0 : first + 8 * (lane - 1), cnt = lane == 0 ? first : 8;
                int brk = -1; uint32_t bcode = RUN_ON;
                if (!resident(plane, qg + 1, c)) brk = before;
                else {
                    const uint2 v = code_word(plane, qg, rev ? rstep(qg, c) : fstep(qg, c));
                    const int w = rev ? W * ((c - 1) / W + 1) - c : (c - 1) % W;
                    const uint32_t codes = ((w < 2 ? v.x : v.y) >> (16 * (w & 1))) & 0xFFFFu;    // 2 bits per row of the group
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int j = dr > 0 ? (lane == 0 ? 8 - first + k : k) : (lane == 0 ? first - 1 - k : 7 - k);
                        const uint32_t code = (codes >> (2 * j)) & 3u;
                        if (k < cnt && brk < 0 && code != t) { brk = before + k; bcode = code; }
                    }
                }
                const unsigned mk = __ballot_sync(0xFFFFFFFFu, brk >= 0);
                if (!mk) { tnext = RUN_ON; return first + 8 * 31; }
                const int src = __ffs(mk) - 1;
                tnext = __shfl_sync(0xFFFFFFFFu, bcode, src);
                return __shfl_sync(0xFFFFFFFFu, brk, src);
            }
        }
#endif
        const int rr = r + lane * dr, cc = c + lane * dc;
        const bool known = !interior(rr, cc) || resident(plane, rr, cc);
        const uint32_t code = known ? code_nc(plane, rr, cc) : RUN_ON;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, code != t);
        if (!m) { tnext = RUN_ON; return 32; }
        const int n = __ffs(m) - 1;
        tnext = __shfl_sync(0xFFFFFFFFu, code, n);
        return n;
    }
    // K7 helper.  Of the cells x0, x0 + dir, ..., x1 (at most 480) of row r of a maxima plane (1 = FM, 3 = RM): which 32-cell groups
    // (bit g: cells x0 + 32 g dir ... x0 + (32 g + 31) dir) hold a cell whose code is not HORIZONTAL.  A HORIZONTAL cell merges into
    // its neighbour's walk whatever that one does (merges_into_next with d = 0), and on the long plateaus of the maxima nearly
    // every cell is one: four trace words per lane classify 480 cells in one pass.  Rows outside the matrix: every group.
    __device__ unsigned groups_not_horizontal(int plane, int r, int x0, int x1, int dir) const {
        const int n = (dir > 0 ? x1 - x0 : x0 - x1) + 1;
        const unsigned all = (1u << ((n + 31) >> 5)) - 1u;
        if (r < 1 || r > len2) return all;
        const int lowcol = max(min(x0, x1), 1), highcol = min(max(x0, x1), len1);
        auto group = [&](int x) { return (dir > 0 ? x - x0 : x0 - x) >> 5; };
        unsigned need = 0;
        if (min(x0, x1) < 1) need |= 1u << group(0);                       // border cells carry no HORIZONTAL code that would merge
        if (max(x0, x1) > len1) need |= 1u << group(len1 + 1);
        const int mlo = (lowcol - 1) >> 2;
        bool ok = true;
#pragma unroll
        for (int g = 0; g < 4; ++g) { const int c1 = 4 * (mlo + 4 * lane + g) + 1; if (c1 <= highcol && !resident(plane, r, c1)) ok = false; }
        if (!__all_sync(0xFFFFFFFFu, ok)) {
#pragma unroll 1
            for (int g = 0; g < 4; ++g) { const int c1 = 4 * (mlo + 4 * lane + g) + 1; ensure_lanes(plane, c1 <= highcol ? r : -1, c1); }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int c1 = 4 * (mlo + 4 * lane + g) + 1;
            if (c1 > highcol) continue;
            const uint32_t codes = (plane >> 1) ? rev_codes(plane, r, c1) : fwd_codes(plane, r, c1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x = c1 + j;
                const uint32_t code = (codes >> (2 * ((plane >> 1) ? 3 - j : j))) & 3u;
                if (x >= lowcol && x <= highcol && code != T_HORI) need |= 1u << group(x);
            }
        }
        return __reduce_or_sync(0xFFFFFFFFu, need);
    }
    // K7 helper.  A maxima-trace walk (plane 1 = FM stepping by -1, plane 3 = RM stepping by +1) that starts at (r, x)
    // merges into the walk that starts at the next cell (r, x + dir) when it makes d >= 0 VERTICAL steps followed by a
    // HORIZONTAL one and the neighbour's walk makes the same d VERTICAL steps first.  Both walks then share their
    // endpoint, so traceBackMaxima + addBpoints of the first is a guaranteed duplicate of the second (:652-655).
    // Evaluated for 32 consecutive start cells at once; lanes with in == false are ignored.
    __device__ bool merges_into_next(int plane, int r, int x, int dir, bool in) const {
        int d0 = 0, d1 = 0;
        uint32_t c0 = T_STOP, c1 = T_STOP;
        bool run0 = in, run1 = in;
        for (int depth = 0; depth < AGE_K7_LADDER; ++depth) {
            ensure_lanes(plane, run0 ? r + depth * dir : -1, x);
            ensure_lanes(plane, run1 ? r + depth * dir : -1, x + dir);
            if (run0) { c0 = code_nc(plane, r + depth * dir, x); if (c0 == T_VERT) d0 = depth + 1; else run0 = false; }
            if (run1) { c1 = code_nc(plane, r + depth * dir, x + dir); if (c1 == T_VERT) d1 = depth + 1; else run1 = false; }
            if (!__any_sync(0xFFFFFFFFu, run0 || run1)) break;
        }
        if (!in || run0 || run1) return false;            // deeper ladders than AGE_K7_LADDER rows are walked one by one
        return c0 == T_HORI && (d0 == 0 || d0 == d1);
    }
    __device__ uint32_t code_at(int plane, int r, int c) const {      // warp-uniform single read
        if (interior(r, c)) ensure(plane, (r - 1) >> 8, win_of(plane, r, c));
        return code_nc(plane, r, c);
    }
    // K6 cell (r, c), 0 <= r <= len2, 0 <= c <= len1: does Fmax[r][c] + Rmax[r+1][c+1] equal the maximum?
    // (interior cells need the RM window of (r+1, c+1) resident)
    __device__ bool tie_nc(int r, int c) const {
        if (r < len2 && c < len1) return hot_nc(r + 1, c + 1);
        return F2(r, c) == max2;                          // Rmax is zero on the border
    }
};

struct BpList {
    int n;
    int (*bp)[4];
};

// K8 traceBackMaxima (:615-642).  The chains are followed run by run: the 32 lanes read the next 32 cells along the
// current direction and the warp jumps to the first cell whose code differs.
template <bool WIDE>
__device__ void walk_chain(const View<WIDE> &V, int plane, int sign, int &c, int &r)
{
    uint32_t t = V.code_at(plane, r, c);
    while (t != T_STOP) {
        const int dr = t == T_VERT ? sign : 0, dc = t == T_VERT ? 0 : sign;
        uint32_t tn;
        const int n = V.run_len(plane, r, c, dr, dc, t, tn);
        r += n * dr; c += n * dc;
        t = tn != View<WIDE>::RUN_ON ? tn : V.code_at(plane, r, c);
    }
}
template <bool WIDE>
__device__ void walk_left(const View<WIDE> &V, int &lg1, int &lg2) { P2_PROF_T0(wc); walk_chain(V, 1, -1, lg1, lg2); P2_OD(V, 6, 1); P2_OD(V, 7, clock64() - wc); }
template <bool WIDE>
__device__ void walk_right(const View<WIDE> &V, int &rg1, int &rg2) { P2_PROF_T0(wc); walk_chain(V, 3, +1, rg1, rg2); P2_OD(V, 6, 1); P2_OD(V, 7, clock64() - wc); }

// K9 addBpoints (:644-668); executed redundantly by every lane on the shared list, lane 0 writes
template <bool WIDE>
__device__ int add_bp(const View<WIDE> &V, BpList &L, int lg1, int lg2, int rg1, int rg2, int lane)
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

// Rows with at most 32 candidate columns each: log2 of the columns tested per row and pass (1, 2, 4, ..., 32), from the widest
// of the 32 rows
__device__ __forceinline__ int narrow_shift(int my_width_minus_1)
{
    const unsigned w = __reduce_max_sync(0xFFFFFFFFu, (unsigned)my_width_minus_1);
    return w == 0 ? 0 : 32 - __clz(w);
}

// K6 findIndelBPs, ordered part (:552-610).  rlo/rhi bound, per row, the columns that can hold a cell with
// Fmax + Rmax == max (from the hot windows of the reverse sweep and the border rule); rows with rhi < rlo are skipped.
template <bool WIDE>
__device__ void indel_enumerate(const View<WIDE> &V, BpList &L, const int *rlo, const int *rhi, int lane)
{
    const int len2 = V.len2;
    int n_add = 0;
    bool stop = false;
    P2_PROF_T0(fc0);
    // Most rows hold no candidate: the per-row column ranges are inspected 32 rows at a time (one per lane) and only the
    // non-empty rows are visited, in the reference's order.
    for (int rb = 0; rb <= len2 && !stop; rb += 32) {                     // forward scan (:552-573)
      const int myr = min(rb + lane, len2), mylo = rlo[myr], myhi = rhi[myr];
      const bool mine = rb + lane <= len2 && myhi >= mylo;
      unsigned rows = __ballot_sync(0xFFFFFFFFu, mine);
      if (rows && __all_sync(0xFFFFFFFFu, !mine || myhi - mylo < 32)) {
        // at most 32 candidate columns per row (the split runs along a column: insertions; a few columns wide when the flank next
        // to it can end in more than one place): the rows are tested side by side, 32 rows x 1 column, 16 x 2, ... or 1 x 32 per
        // pass; lane order = row-major order, hits are taken in that order
        const int sh = narrow_shift(mine ? myhi - mylo : 0), rp = 32 >> sh;
        for (int sub = 0; sub < 32 && !stop; sub += rp) {
            if (!((rows >> sub) & (0xFFFFFFFFu >> (32 - rp)))) continue;
            P2_OD(V, 4, 1);
            const int src = sub + (lane >> sh);
            const int r = rb + src, lo = __shfl_sync(0xFFFFFFFFu, mylo, src), hi = __shfl_sync(0xFFFFFFFFu, myhi, src);
            const int c = lo + (lane & ((1 << sh) - 1));
            const bool in = ((rows >> src) & 1u) && c <= hi;
            V.ensure_lanes(3, in ? r + 1 : -1, c + 1);
            const bool tie = in && V.tie_nc(r, c);
            V.ensure_lanes(1, tie ? r : -1, c);
            unsigned m = __ballot_sync(0xFFFFFFFFu, tie && V.code_nc(1, r, c) == T_STOP);
            while (m && !stop) {
                const int b = __ffs(m) - 1; m &= m - 1;
                const int cc = __shfl_sync(0xFFFFFFFFu, c, b), rr = rb + sub + (b >> sh);
                int lg1 = cc, lg2 = rr, rg1 = cc + 1, rg2 = rr + 1;
                walk_left(V, lg1, lg2); walk_right(V, rg1, rg2);
                const int res = add_bp(V, L, lg1, lg2, rg1, rg2, lane);
                if (res > 0) ++n_add;
                if (n_add >= MAX_BP / 2 || res < 0) stop = true;
            }
        }
        continue;
      }
      while (rows && !stop) {
        const int r = rb + __ffs(rows) - 1; rows &= rows - 1;
        const int lo = rlo[r], hi = rhi[r];
        P2_OD(V, 5, 1);
        auto candidate = [&](int c) {                                   // cell (r, c) ties at the maximum and starts a maxima chain
            int lg1 = c, lg2 = r, rg1 = c + 1, rg2 = r + 1;
            walk_left(V, lg1, lg2); walk_right(V, rg1, rg2);
            const int res = add_bp(V, L, lg1, lg2, rg1, rg2, lane);
            if (res > 0) ++n_add;
            if (n_add >= MAX_BP / 2 || res < 0) stop = true;
        };
        auto slow = [&](int a, int b) {                                 // 32 cells per pass; any cell, borders included
            for (int cb = a & ~31; cb <= b && !stop; cb += 32) {
                const int c = cb + lane;
                const bool in = c >= a && c <= b;
                V.ensure_lanes(3, in ? r + 1 : -1, c + 1);
                const bool tie = in && V.tie_nc(r, c);
                V.ensure_lanes(1, tie ? r : -1, c);
                unsigned m = __ballot_sync(0xFFFFFFFFu, tie && V.code_nc(1, r, c) == T_STOP);
                while (m && !stop) { const int b2 = __ffs(m) - 1; m &= m - 1; candidate(cb + b2); }
            }
        };
        // Long runs along a row (a deletion: every column of the excised stretch ties): interior cells 1 <= c <= len1 - 1 of
        // interior rows go through 128 cells per pass -- one forward block per lane: its FM word, and the hot words of the
        // two reverse blocks that hold the reverse cells (r+1, c+1)
        const int a = max(lo, 1), b = min(hi, V.len1 - 1);
        if (r >= 1 && r <= len2 - 1 && b - a >= 96) {
            if (lo < a) slow(lo, a - 1);
            for (int g0 = (a - 1) >> 2; g0 <= (b - 1) >> 2 && !stop; g0 += 32) {
                const int m = g0 + lane, c1 = 4 * m + 1;
                const bool act = m <= (b - 1) >> 2, fourth = act && c1 + 4 <= V.len1 && c1 + 3 >= a && c1 + 3 <= b;      // (its hot word is only read for cell c1 + 3)
                V.ensure_lanes(1, act ? r : -1, c1);
                V.ensure_lanes(3, act ? r + 1 : -1, c1 + 1);
                V.ensure_lanes(3, fourth ? r + 1 : -1, c1 + 4);
                uint32_t hits = 0;
                if (act) {
                    const uint32_t fm = V.fwd_codes(1, r, c1), hA = V.hot_word(r + 1, c1 + 1), hB = fourth ? V.hot_word(r + 1, c1 + 4) : 0u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = c1 + j;
                        const uint32_t tie = j < 3 ? (hA >> (8 * (2 - j) + (r & 7))) & 1u : (hB >> (24 + (r & 7))) & 1u;
                        if (c >= a && c <= b && tie && ((fm >> (2 * j)) & 3u) == T_STOP) hits |= 1u << j;
                    }
                }
                unsigned lanes = __ballot_sync(0xFFFFFFFFu, hits != 0);
                while (lanes && !stop) {
                    const int src = __ffs(lanes) - 1; lanes &= lanes - 1;
                    uint32_t h = __shfl_sync(0xFFFFFFFFu, hits, src);
                    while (h && !stop) { const int j = __ffs(h) - 1; h &= h - 1; candidate(4 * (g0 + src) + 1 + j); }
                }
            }
            if (hi > b && !stop) slow(b + 1, hi);
        } else slow(lo, hi);
      }
    }
#ifdef AGE_P2_PROF
    if (lane == 0 && V.od) V.od[3] += (unsigned int)(clock64() - fc0);
#endif
    n_add = 0; stop = false;
    for (int rb = len2 & ~31; rb >= 0 && !stop; rb -= 32) {               // reverse scan (:586-610)
      const int myr = min(rb + lane, len2), mylo = rlo[myr], myhi = rhi[myr];
      const bool mine = rb + lane <= len2 && myhi >= mylo;
      unsigned rows = __ballot_sync(0xFFFFFFFFu, mine);
      if (rows && __all_sync(0xFFFFFFFFu, !mine || myhi - mylo < 32)) {
        const int sh = narrow_shift(mine ? myhi - mylo : 0), rp = 32 >> sh;
        for (int sub = 32 - rp; sub >= 0 && !stop; sub -= rp) {
            if (!((rows >> sub) & (0xFFFFFFFFu >> (32 - rp)))) continue;
            P2_OD(V, 4, 1);
            const int src = sub + (lane >> sh);
            const int r = rb + src, lo = __shfl_sync(0xFFFFFFFFu, mylo, src), hi = __shfl_sync(0xFFFFFFFFu, myhi, src);
            const int c = lo + (lane & ((1 << sh) - 1));
            const bool in = ((rows >> src) & 1u) && c <= hi;
            V.ensure_lanes(3, in ? r + 1 : -1, c + 1);
            unsigned m = __ballot_sync(0xFFFFFFFFu, in && V.tie_nc(r, c) && V.code_nc(3, r + 1, c + 1) == T_STOP);
            while (m && !stop) {
                const int b = 31 - __clz(m); m &= ~(1u << b);
                const int cc = __shfl_sync(0xFFFFFFFFu, c, b), rr = rb + sub + (b >> sh);
                int lg1 = cc, lg2 = rr, rg1 = cc + 1, rg2 = rr + 1;
                walk_left(V, lg1, lg2); walk_right(V, rg1, rg2);
                const int res = add_bp(V, L, lg1, lg2, rg1, rg2, lane);
                if (res > 0) ++n_add;
                if (n_add >= MAX_BP / 2 || res < 0) stop = true;
            }
        }
        continue;
      }
      while (rows && !stop) {
        const int r = rb + 31 - __clz(rows); rows &= ~(1u << (r - rb));
        const int lo = rlo[r], hi = rhi[r];
        P2_OD(V, 5, 1);
        auto candidate = [&](int c) {
            int lg1 = c, lg2 = r, rg1 = c + 1, rg2 = r + 1;
            walk_left(V, lg1, lg2); walk_right(V, rg1, rg2);
            const int res = add_bp(V, L, lg1, lg2, rg1, rg2, lane);
            if (res > 0) ++n_add;
            if (n_add >= MAX_BP / 2 || res < 0) stop = true;
        };
        auto slow = [&](int a, int b) {
            for (int cb = b & ~31; cb >= (a & ~31) && !stop; cb -= 32) {
                const int c = cb + lane;
                const bool in = c >= a && c <= b;
                V.ensure_lanes(3, in ? r + 1 : -1, c + 1);
                unsigned m = __ballot_sync(0xFFFFFFFFu, in && V.tie_nc(r, c) && V.code_nc(3, r + 1, c + 1) == T_STOP);
                while (m && !stop) { const int b2 = 31 - __clz(m); m &= ~(1u << b2); candidate(cb + b2); }
            }
        };
        // the same long runs from the right: one reverse block per lane (reverse cells (r+1, c+1) with c + 1 = 4m+1 .. 4m+4),
        // its RM word and its hot word, the block with the largest columns first
        const int a = lo, b = min(hi, V.len1 - 1);
        if (r <= len2 - 1 && b - a >= 96) {
            if (hi > b) slow(b + 1, hi);
            for (int g0 = b >> 2; g0 >= (a >> 2) && !stop; g0 -= 32) {
                const int m = g0 - lane, c1 = 4 * m + 1;                // reverse cells (r+1, c1 .. c1+3) = cells (r, c1-1 .. c1+2)
                const bool act = m >= (a >> 2);
                V.ensure_lanes(3, act ? r + 1 : -1, c1);
                uint32_t hits = 0;
                if (act) {
                    const uint32_t rm = V.rev_codes(3, r + 1, c1), hw = V.hot_word(r + 1, c1);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = c1 + j - 1;
                        if (c >= a && c <= b && ((hw >> (8 * (3 - j) + (r & 7))) & 1u) && ((rm >> (2 * (3 - j))) & 3u) == T_STOP) hits |= 1u << j;
                    }
                }
                unsigned lanes = __ballot_sync(0xFFFFFFFFu, hits != 0);
                while (lanes && !stop) {
                    const int src = __ffs(lanes) - 1; lanes &= lanes - 1;
                    uint32_t h = __shfl_sync(0xFFFFFFFFu, hits, src);
                    while (h && !stop) { const int j = 31 - __clz(h); h &= ~(1u << j); candidate(4 * (g0 - src) + j); }
                }
            }
        } else slow(lo, hi);
      }
    }
}

// Hot-window refinement: every reverse window that holds a (lane, window) tile whose maximum equals the split
// maximum is re-swept with the trace on; its hot bits widen the column range of the K6 rows they belong to.
template <bool WIDE>
__device__ void indel_row_ranges(const View<WIDE> &V, int *rlo, int *rhi, int lane)
{
    const PairTask &T = *V.T;
    const int len1 = V.len1, len2 = V.len2, nb = V.nb, max2 = V.max2;
    for (int r = lane; r <= len2; r += 32) { rlo[r] = 0x7FFFFFFF; rhi[r] = -1; }
    __syncwarp();
    if (max2 == 0) {                                                  // everything ties at zero: scan it all
        for (int r = lane; r <= len2; r += 32) { rlo[r] = 0; rhi[r] = len1; }
        __syncwarp();
        return;
    }
    const int nhot = V.sc.hotlist[0];                                 // written by the select kernel
    // Scratch in the warp's shared memory (the rings and the trace staging are only used while a window is re-swept): the
    // hot words of the window [step][lane] and the column range of each of the lane's rows.  The scan below is a pair of
    // short rolled loops: unrolled over the 32 words x 4 columns x 8 rows it was 3 400 instructions that every hot window
    // (hundreds per long deletion) streamed through the instruction cache.
    uint32_t *hs = reinterpret_cast<uint32_t *>(V.R);                 // [CKS][32]
    int *lo8 = reinterpret_cast<int *>(V.X), *hi8 = lo8 + RPT * 32;   // [RPT][32]
    static_assert(sizeof(WarpRings) >= CKS * 32 * 4 && sizeof(TraceSmem) >= 2 * RPT * 32 * 4, "scratch of indel_row_ranges");
    for (int hi = 0; hi < nhot; ++hi) {
        const int wi = V.sc.hotlist[1 + hi];
        const int s = wi / T.nwin, win = wi % T.nwin;
        V.ensure(3, s, win);
        __syncwarp();
        const int t0 = CKS * win + 1, t1 = min(CKS * win + CKS, T.nsteps);
        const uint32_t *hotpage = reinterpret_cast<const uint32_t *>(V.page(3, s * STRIP, t0) + 2 * P2_GRANULE);   // [step in window][lane]
#pragma unroll
        for (int i = 0; i < CKS; ++i) {                               // the lane's hot words of all steps of the window, in flight together
            const int t = t0 + i, bq = t - 31 + lane;                 // logical reverse block of this lane in step t
            hs[i * 32 + lane] = (t <= t1 && bq >= 1 && bq <= nb) ? __ldcg(hotpage + i * 32 + lane) : 0u;
        }
#pragma unroll
        for (int k = 0; k < RPT; ++k) { lo8[k * 32 + lane] = 0x7FFFFFFF; hi8[k * 32 + lane] = -1; }
        // The hot cells of a lane lie in its own 8 rows.  The columns of step i are cb - 3 .. cb with cb = cb0 - 4 i: the first step
        // that shows a row holds the row's largest column, the last one its smallest.
        const int cb0 = W * (nb + 1 - (t0 - 31 + lane));              // reverse cell column of block column w in step i: cb0 - W i - w
        auto word = [&](int i, int &cb) {                             // hot word of step i without the padding columns behind len1
            const uint32_t x = hs[i * 32 + lane];
            cb = cb0 - W * i;
            const int pad = min(max(cb - len1, 0), W);
            return pad >= W ? 0u : x & (0xFFFFFFFFu << (8 * pad));
        };
        uint32_t seen = 0;
#pragma unroll 1
        for (int i = 0; i < CKS; ++i) {
            int cb;
            const uint32_t x = word(i, cb);
            uint32_t nw = ((x | (x >> 8) | (x >> 16) | (x >> 24)) & 0xFFu) & ~seen;     // rows seen for the first time
            seen |= nw;
            while (nw) {
                const int k = __ffs(nw) - 1; nw &= nw - 1;
                const uint32_t mk = (x >> k) & 0x01010101u;           // bit 8 w: cell (row k, block column w) is hot
                hi8[k * 32 + lane] = cb - ((__ffs(mk) - 1) >> 3) - 1;
            }
        }
        seen = 0;
#pragma unroll 1
        for (int i = CKS - 1; i >= 0; --i) {
            int cb;
            const uint32_t x = word(i, cb);
            uint32_t nw = ((x | (x >> 8) | (x >> 16) | (x >> 24)) & 0xFFu) & ~seen;
            seen |= nw;
            while (nw) {
                const int k = __ffs(nw) - 1; nw &= nw - 1;
                const uint32_t mk = (x >> k) & 0x01010101u;
                lo8[k * 32 + lane] = cb - ((31 - __clz(mk)) >> 3) - 1;
            }
        }
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = s * STRIP + lane * RPT + k + 1;             // reverse cell row; K6 row r - 1
            const int h8 = hi8[k * 32 + lane];
            if (h8 >= 0 && r <= len2) { rlo[r - 1] = min(rlo[r - 1], lo8[k * 32 + lane]); rhi[r - 1] = max(rhi[r - 1], h8); }
        }
    }
    __syncwarp();
    // border cells of K6 (r = len2 or c = len1): Rmax is zero there, so they tie only when Fmax itself == max
    if (V.F2(len2, len1) == max2) {
        int lo = 1, hi = len1;                                        // Fmax[len2][c] is non-decreasing in c
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (V.F2(len2, mid) >= max2) hi = mid; else lo = mid + 1; }
        if (lane == 0) { rlo[len2] = min(rlo[len2], lo); rhi[len2] = len1; }
        for (int r = 1 + lane; r < len2; r += 32)
            if (V.F2(r, len1) == max2) { rlo[r] = min(rlo[r], len1); rhi[r] = len1; }
    }
    __syncwarp();
}

// Smallest x in [lo, hi) with pred(x), or hi when there is none; pred is monotone (false ... false true ... true).  The 32
// lanes probe 32 positions per round (each probe is a dependent chain of global loads: a 20 000-column row takes 3 rounds
// instead of the 15 steps of a bisection).  Warp-collective: every lane calls it with the same arguments.
template <class Pred>
__device__ __forceinline__ int first_true32(int lo, int hi, int lane, Pred pred)
{
    while (hi - lo > 32) {
        const int step = (hi - lo + 31) / 32;
        const int x = min(lo + (lane + 1) * step - 1, hi - 1);
        const unsigned m = __ballot_sync(0xFFFFFFFFu, pred(x));
        if (!m) return hi;                                  // the last probe is hi - 1
        const int k = __ffs(m) - 1;
        hi = min(lo + (k + 1) * step - 1, hi - 1);          // that probe holds: the answer is it unless an earlier position does
        lo = lo + k * step;                                 // the probe before it does not
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, lo + lane < hi && pred(lo + lane));
    return m ? lo + __ffs(m) - 1 : hi;
}

// K7 findInversionTDuplicationBPs (:472-538).  The reference walks every (j1, j2) pair; here the monotone
// rows are binary-searched for the run of matching values, and of the start cells whose traceBackMaxima walks
// merge (View::merges_into_next: they share the endpoint, hence the addBpoints key) only the last one is walked --
// the others are guaranteed duplicates (:652-655) and change nothing.
template <bool WIDE>
__device__ void invtdup_enumerate(const View<WIDE> &V, BpList &L, int lane)
{
    const int len1 = V.len1, len2 = V.len2, max2 = V.max2;
    const uint32_t *colF = V.T->Dlast, *colR = V.T->Rfirst;
    auto A = [&](int r) { return r >= 1 ? V.pick(colF[r]) : 0; };                 // 2*Fmax[r][len1]
    auto B = [&](int r) { return (r >= 1 && r <= len2) ? V.pick(colR[r]) : 0; };  // 2*Rmax[r][1]
    int n_add = 0;
    bool stop = false;
    P2_PROF_T0(kc);
    // one start cell (r, j1) of the forward scan: FM is 0 there
    auto fwd_start = [&](int r, int j1) {
        const int tv = max2 - V.F2(r, j1);
        if (tv < 0) return;
        // columns x in [1, len1+1] of row r+1, values non-increasing: first x with value <= tv
        const int lo = first_true32(1, len1 + 2, lane, [&](int x) { return V.R2(r + 1, x) <= tv; });
        if (V.R2(r + 1, lo) != tv) return;
        // the run of columns with value tv ends before the first x with value < tv
        const int xhi = first_true32(lo, len1 + 2, lane, [&](int x) { return V.R2(r + 1, x) < tv; }) - 1;
        for (int xb0 = lo; xb0 <= xhi && !stop; xb0 += 480) {
          unsigned need = V.groups_not_horizontal(3, r + 1, xb0, min(xhi, xb0 + 479), +1);
          while (need && !stop) {                                  // the 32-cell groups that hold anything but HORIZONTAL codes
            const int xb = xb0 + 32 * (__ffs(need) - 1); need &= need - 1;
            const int xi = xb + lane;
            const bool in = xi <= xhi;
            const bool mg = V.merges_into_next(3, r + 1, xi, +1, in);      // warp-collective: every lane calls it
            unsigned am = __ballot_sync(0xFFFFFFFFu, in && !mg);
            while (am && !stop) {                                  // walks that do not merge into their neighbour's
                const int x = xb + __ffs(am) - 1; am &= am - 1;
                int rg1 = x, rg2 = r + 1, lg1 = j1, lg2 = r;
                walk_right(V, rg1, rg2);
                const int res = add_bp(V, L, lg1, lg2, rg1, rg2, lane);
                if (res > 0) ++n_add;
                if (n_add >= MAX_BP / 2 || res < 0) stop = true;
            }
          }
        }
    };
    for (int rb = 0; rb <= len2 && !stop; rb += 32) {                     // forward (:485-504): the tie rows, 32 rows per look
      unsigned tie_rows = __ballot_sync(0xFFFFFFFFu, rb + lane <= len2 && A(rb + lane) + B(rb + lane + 1) == max2);
      while (tie_rows && !stop) {
        const int r = rb + __ffs(tie_rows) - 1; tie_rows &= tie_rows - 1;
        P2_PROF2_ADD(0, 1);
        P2_OD(V, 4, 1);
        P2_PROF_T0(ec);
        if (r + 1 <= len2) V.ensure_rm(r >> 8);                           // row r+1 of Rmax and RM
        P2_PROF2_ADD(5, clock64() - ec);
        P2_OD(V, 3, clock64() - ec);
        if (r == 0) {                                                     // border row: only (0, 0) has FM = 0 (:874-882)
            fwd_start(0, 0);
            continue;
        }
        // interior row: column 0 carries FM = VERTICAL; columns 1 .. len1 four forward blocks per lane, 512 cells per pass (a pass
        // is one chain of dependent loads -- window bit, page table, trace word -- whatever the number of words it fetches)
        for (int g0 = 0; g0 <= (len1 - 1) >> 2 && !stop; g0 += 128) {
            bool ok = true;
#pragma unroll
            for (int g = 0; g < 4; ++g) { const int c1 = 4 * (g0 + 4 * lane + g) + 1; if (c1 <= len1 && !V.resident(1, r, c1)) ok = false; }
            if (!__all_sync(0xFFFFFFFFu, ok)) {
#pragma unroll 1
                for (int g = 0; g < 4; ++g) { const int c1 = 4 * (g0 + 4 * lane + g) + 1; V.ensure_lanes(1, c1 <= len1 ? r : -1, c1); }
            }
            uint32_t hits = 0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int c1 = 4 * (g0 + 4 * lane + g) + 1;
                if (c1 <= len1) {
                    const uint32_t fm = V.fwd_codes(1, r, c1);
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (c1 + j <= len1 && ((fm >> (2 * j)) & 3u) == T_STOP) hits |= 1u << (4 * g + j);
                }
            }
            unsigned lanes = __ballot_sync(0xFFFFFFFFu, hits != 0);
            while (lanes && !stop) {
                const int src = __ffs(lanes) - 1; lanes &= lanes - 1;
                uint32_t h = __shfl_sync(0xFFFFFFFFu, hits, src);
                while (h && !stop) {
                    const int b = __ffs(h) - 1; h &= h - 1;
                    P2_PROF_T0(sc0);
                    fwd_start(r, 4 * (g0 + 4 * src + (b >> 2)) + 1 + (b & 3));
                    P2_PROF2_ADD(1, 1); P2_PROF2_ADD(3, clock64() - sc0); P2_OD(V, 5, 1); P2_OD(V, 1, clock64() - sc0);
                }
            }
        }
      }
    }
    P2_PROF_ADD(1, kc);
    P2_PROF_RESET(kc);
    n_add = 0; stop = false;
    // one start cell (i2, j2) of the reverse scan: RM is 0 there
    auto rev_start = [&](int i2, int j2) {
        const int tv = max2 - V.R2(i2, j2);
        if (tv < 0) return;
        // columns x in [0, len1] of row i2-1, values non-decreasing: last x with value <= tv
        const int lo = first_true32(0, len1 + 1, lane, [&](int x) { return V.F2(i2 - 1, x) > tv; }) - 1;
        if (lo < 0 || V.F2(i2 - 1, lo) != tv) return;
        // the run of columns with value tv starts after the last x with value < tv
        const int xlo = first_true32(0, lo + 1, lane, [&](int x) { return V.F2(i2 - 1, x) >= tv; });
        for (int xb0 = lo; xb0 >= xlo && !stop; xb0 -= 480) {
          unsigned need = V.groups_not_horizontal(1, i2 - 1, xb0, max(xlo, xb0 - 479), -1);
          while (need && !stop) {
            const int xb = xb0 - 32 * (__ffs(need) - 1); need &= need - 1;
            const int xi = xb - lane;
            const bool in = xi >= xlo;
            const bool mg = V.merges_into_next(1, i2 - 1, xi, -1, in);
            unsigned am = __ballot_sync(0xFFFFFFFFu, in && !mg);
            while (am && !stop) {
                const int x = xb - (__ffs(am) - 1); am &= am - 1;
                int lg1 = x, lg2 = i2 - 1, rg1 = j2, rg2 = i2;
                walk_left(V, lg1, lg2);
                const int res = add_bp(V, L, lg1, lg2, rg1, rg2, lane);
                if (res > 0) ++n_add;
                if (n_add >= MAX_BP / 2 || res < 0) stop = true;
            }
          }
        }
    };
    for (int rb = (len2 + 1) & ~31; rb >= 0 && !stop; rb -= 32) {          // reverse (:515-534): the tie rows from the last one
      unsigned tie_rows = __ballot_sync(0xFFFFFFFFu, rb + lane >= 1 && rb + lane <= len2 + 1 && A(rb + lane - 1) + B(rb + lane) == max2);
      while (tie_rows && !stop) {
        const int i2 = rb + 31 - __clz(tie_rows); tie_rows &= ~(1u << (i2 - rb));
        P2_OD(V, 4, 1);
        P2_PROF_T0(ec2);
        if (i2 <= len2) V.ensure_rm((i2 - 1) >> 8);
        P2_OD(V, 3, clock64() - ec2);
        if (i2 > len2) {                                                  // border row: RM is 0 in every column (:924-932)
            for (int j2 = len1 + 1; j2 >= 1 && !stop; --j2) { P2_PROF_T0(sc0); rev_start(i2, j2); P2_OD(V, 5, 1); P2_OD(V, 1, clock64() - sc0); }
            continue;
        }
        rev_start(i2, len1 + 1);                                          // border column
        // columns len1 .. 1, four reverse blocks per lane, the block with the largest columns first
        for (int g0 = (len1 - 1) >> 2; g0 >= 0 && !stop; g0 -= 128) {
            bool ok = true;
#pragma unroll
            for (int g = 0; g < 4; ++g) { const int m = g0 - (4 * lane + g); if (m >= 0 && !V.resident(3, i2, 4 * m + 1)) ok = false; }
            if (!__all_sync(0xFFFFFFFFu, ok)) {
#pragma unroll 1
                for (int g = 0; g < 4; ++g) { const int m = g0 - (4 * lane + g); V.ensure_lanes(3, m >= 0 ? i2 : -1, 4 * m + 1); }
            }
            uint32_t hits = 0;                                            // bit 4 * (3 - g) + j: column offset j of the lane's g-th block
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int m = g0 - (4 * lane + g), c1 = 4 * m + 1;
                if (m >= 0) {
                    const uint32_t rm = V.rev_codes(3, i2, c1);
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (c1 + j <= len1 && ((rm >> (2 * (3 - j))) & 3u) == T_STOP) hits |= 1u << (4 * (3 - g) + j);
                }
            }
            unsigned lanes = __ballot_sync(0xFFFFFFFFu, hits != 0);
            while (lanes && !stop) {
                const int src = __ffs(lanes) - 1; lanes &= lanes - 1;
                uint32_t h = __shfl_sync(0xFFFFFFFFu, hits, src);
                while (h && !stop) {
                    const int b = 31 - __clz(h); h &= ~(1u << b);
                    P2_PROF_T0(sc0);
                    rev_start(i2, 4 * (g0 - 4 * src - (3 - (b >> 2))) + 1 + (b & 3));
                    P2_PROF2_ADD(2, 1); P2_PROF2_ADD(4, clock64() - sc0); P2_OD(V, 5, 1); P2_OD(V, 1, clock64() - sc0);
                }
            }
        }
      }
    }
    P2_PROF_ADD(2, kc);
}

// K11 findLeftBlocks (:752-808).  dst == nullptr: count only -- the blocks found are left in `buf` (shared memory, walking
// order, at most `cap` of them) so that the caller need not walk the path a second time once it has room for them.  Blocks are
// produced last-first; a block is a maximal run of DIAGONAL codes (any H / V step in between starts a new block, :774-790).
template <bool WIDE>
__device__ int left_blocks(const View<WIDE> &V, int c, int r, Block *dst, int n_total, int lane, Block *buf = nullptr, int cap = 0)
{
    int n = 0, run = 0;
    auto put = [&](const Block &b) {
        if (lane == 0) {
            if (dst && n < n_total) dst[n_total - 1 - n] = b;
            if (buf && n < cap) buf[n] = b;
        }
        ++n;
    };
    uint32_t t = V.code_at(0, r, c);
    while (t != T_STOP) {
        const int dr = t == T_HORI ? 0 : -1, dc = t == T_VERT ? 0 : -1;
        uint32_t tn;
        const int len = V.run_len(0, r, c, dr, dc, t, tn);
        r += len * dr; c += len * dc;
        if (t == T_DIAG) run += len;
        else if (run > 0) { put(Block{c - len * dc + 1, r - len * dr + 1, run}); run = 0; }
        t = tn != View<WIDE>::RUN_ON ? tn : V.code_at(0, r, c);
    }
    if (run > 0) put(Block{c + 1, r + 1, run});
    return n;       // (n <= n_total unless the trace pool overflowed between the counting and the writing pass: the batch is run again then)
}

// K12 findRightBlocks (:810-860)
template <bool WIDE>
__device__ int right_blocks(const View<WIDE> &V, int c, int r, Block *dst, int n_total, int lane, Block *buf = nullptr, int cap = 0)
{
    int n = 0, run = 0, s1 = 0, s2 = 0;
    auto put = [&](const Block &b) {
        if (lane == 0) {
            if (dst && n < n_total) dst[n] = b;
            if (buf && n < cap) buf[n] = b;
        }
        ++n;
    };
    uint32_t t = V.code_at(2, r, c);
    while (t != T_STOP) {
        const int dr = t == T_HORI ? 0 : 1, dc = t == T_VERT ? 0 : 1;
        uint32_t tn;
        const int len = V.run_len(2, r, c, dr, dc, t, tn);
        if (t == T_DIAG) { if (run == 0) { s1 = c; s2 = r; } run += len; }     // interior cells only (:820 always holds)
        else if (run > 0) { put(Block{s1, s2, run}); run = 0; }
        r += len * dr; c += len * dc;
        t = tn != View<WIDE>::RUN_ON ? tn : V.code_at(2, r, c);
    }
    if (run > 0) put(Block{s1, s2, run});
    return n;
}

__device__ __forceinline__ P2Scratch carve_set(const PairTask &T, int set, char *tpool, uint32_t tpool_granules, uint32_t *tpool_next)
{
    const size_t nwt = (size_t)T.nstrips * T.nwin;
    char *base = T.p2set + (size_t)set * T.p2stride;
    P2Scratch sc;
    sc.ptab = (int32_t *)base; base += p2_up(5 * nwt * 4);
    sc.pool = tpool; sc.pool_granules = tpool_granules; sc.pool_next = tpool_next;
    sc.vwords = (int)p2_vwords(nwt); sc.nwt = (int)nwt;
    sc.valid = (uint32_t *)base; base += p2_up((size_t)9 * sc.vwords * 4);
    sc.rowrange = (int32_t *)base; base += p2_up(2 * ((size_t)T.len2 + 2) * 4);
    sc.hotlist = (int32_t *)base;
    return sc;
}
__device__ __forceinline__ P2Scratch carve_set(const PairTask &T, int set, const Pass2Params &prm)
{
    return carve_set(T, set, prm.tpool, prm.tpool_granules, prm.tpool_next);
}

// Split maxima of both halves (:545-550 / :476-481) and the -both / -inv selection (age_align.cpp:273 /
// AGEaligner.cpp:151: ties keep half 0).  The half that loses its selection gets its score-only record here.
template <bool WIDE>
__device__ void pair_select(const PairTask &T, AlignerOut *outs, int lane, int (&max2)[2], bool (&need)[2])
{
    using A = Ar<WIDE>;
    const int len2 = T.len2;
    uint32_t mx_indel = 0, mx_col = 0;                       // max of 2*(Fmax + Rmax), packed per half
    const int nwt = T.nstrips * T.nwin;
    {   // nwt * 32 words = nwt * 8 uint4: four independent 16-byte loads per lane and pass
        const uint4 *tm = reinterpret_cast<const uint4 *>(T.tilemax);
        const int n4 = nwt * 8;
        for (int i = lane; i < n4; i += 128) {
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (i + 32 * k < n4) ? __ldg(tm + i + 32 * k) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) mx_indel = A::maxu(A::maxu(mx_indel, A::maxu(v[k].x, v[k].y)), A::maxu(v[k].z, v[k].w));
        }
    }
    for (int r = lane; r <= len2; r += 32) {
        const uint32_t f = r >= 1 ? T.Dlast[r] : 0u, q = r + 1 <= len2 ? T.Rfirst[r + 1] : 0u;
        mx_col = A::maxu(mx_col, A::sum(A::add(f, q)));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mx_indel = A::maxu(mx_indel, __shfl_xor_sync(0xFFFFFFFFu, mx_indel, o));
        mx_col   = A::maxu(mx_col,   __shfl_xor_sync(0xFFFFFFFFu, mx_col, o));
    }
    mx_indel = A::maxu(mx_indel, T.Dlast[len2]);             // Fmax[len2][len1] + 0 (border cells)
    for (int h = 0; h < 2; ++h) {
        const uint32_t v = (T.flag[h] == F_INDEL) ? mx_indel : mx_col;
        max2[h] = WIDE ? (int)v : (int)((v >> (16 * h)) & 0xFFFFu);
        need[h] = T.out[h] >= 0 && T.flag[h] != 0;
    }
    if (T.select && need[0] && need[1]) {
        const int win = max2[1] > max2[0] ? 1 : 0;
        if (lane == 0) {
            AlignerOut &Lo = outs[T.out[1 - win]];
            Lo.status = 0; Lo.score = max2[1 - win] >> 1; Lo.n_bp = -1; Lo.alt_index = -1;
            for (int q = 0; q < 4; ++q) { Lo.n_blk[q] = 0; Lo.blk_off[q] = 0; }
        }
        need[1 - win] = false;
    }
}

// INDEL: the reverse windows that hold a (lane, window) tile at the split maximum (sc.hotlist), in window order.
// 32 windows per pass: every lane looks through the 32 tile words of one window (eight 16-byte loads in flight), a
// ballot collects the windows that hold the maximum.
template <bool WIDE>
__device__ int build_hotlist(const PairTask &T, const P2Scratch &sc, int h, int max2h, int lane)
{
    const int nwt = T.nstrips * T.nwin;
    const uint4 *tm = reinterpret_cast<const uint4 *>(T.tilemax);
    int nhot = 0;
    for (int base = 0; base < nwt; base += 32) {
        const int wi = base + lane;
        bool hot = false;
        if (wi < nwt) {
            uint4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldg(tm + (size_t)wi * 8 + i);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t w4[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) hot = hot || (WIDE ? (int)w4[j] : (int)((w4[j] >> (16 * h)) & 0xFFFFu)) == max2h;
            }
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, hot);
        if (hot) sc.hotlist[1 + nhot + __popc(m & ((1u << lane) - 1u))] = wi;
        nhot += __popc(m);
    }
    if (lane == 0) sc.hotlist[0] = nhot;
    __syncwarp();
    return nhot;
}

// K5-K9 of one aligner: fills L (shared memory) and the breakpoint part of its output record
template <bool WIDE>
__device__ void enumerate_aligner(const View<WIDE> &V, int flag, BpList &L, AlignerOut &O, int lane)
{
    __syncwarp();
    if (lane < 4) L.bp[0][lane] = 0;        // no breakpoint at all (only possible when the u16 split sum wrapped): findAlignment(0) of zeros
    __syncwarp();
    if (flag == F_INDEL) {
        int *rlo = V.sc.rowrange, *rhi = V.sc.rowrange + (V.len2 + 2);
        P2_PROF_T0(pc);
        indel_row_ranges(V, rlo, rhi, lane);
#ifdef AGE_P2_PROF
        if (lane == 0 && V.od) { V.od[1] += (unsigned int)(clock64() - pc); V.od[2] += (unsigned int)V.sc.hotlist[0]; }
#endif
        P2_PROF_ADD(1, pc);
        P2_PROF_RESET(pc);
        indel_enumerate(V, L, rlo, rhi, lane);
        P2_PROF_ADD(2, pc);
    } else invtdup_enumerate(V, L, lane);
    __syncwarp();
    if (lane == 0) { O.status = 0; O.score = V.max2 >> 1; O.n_bp = L.n; }
    for (int i = lane; i < L.n * 4; i += 32) O.bp[i >> 2][i & 3] = L.bp[i >> 2][i & 3];
}

// ---- numbers of the printed alignment (AlignerOut::call) ----------------------------------------------------------------
__device__ __forceinline__ int nuc_up(char c)               // upper-case nucleotide, 0 for everything Sequence::sameNuc rejects (Sequence.hh:182-188)
{
    const int u = (c >= 'a' && c <= 'z') ? c - 32 : c;
    return (u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'U') ? u : 0;
}
__device__ __forceinline__ bool same_nuc_d(char a, char b) { const int x = nuc_up(a); return x != 0 && x == nuc_up(b); }
__device__ __forceinline__ bool is_gap_d(char c) { return c == ' ' || c == '-'; }           // Sequence.hh:175-180
__device__ __forceinline__ char complement_d(char c)        // Sequence.hh:138-150
{
    switch (c) {
    case 'a': return 't'; case 'c': return 'g'; case 't': return 'a'; case 'g': return 'c';
    case 'A': return 'T'; case 'C': return 'G'; case 'T': return 'A'; case 'G': return 'C';
    case 'u': return 'a'; case 'U': return 'A';
    default:  return c;
    }
}
struct SeqD {                                               // a sequence as passed (0-based positions; '\0' outside)
    const char *s; int n;
    __device__ __forceinline__ char at(int i) const { return (i >= 0 && i < n) ? s[i] : '\0'; }
};
__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// AliFragment::countAligned of one fragment from its diagonal blocks: alignment columns, identical columns, gap columns.
// rc1: the fragment's first-sequence positions are positions of the reverse complement (right fragment of INVL, left of INVR).
__device__ void fragment_counts(const Block *blk, int n, const SeqD &q1, const SeqD &q2, bool rc1, int lane, int &n_ali, int &n_id, int &n_gap)
{
    int al = 0, id = 0, gp = 0;
    for (int b = 0; b < n; ++b) {
        const Block B = blk[b];
        if (b + 1 < n && lane == 0) {                       // the columns between two blocks pair a base with a gap
            const int g = (blk[b + 1].start1 - (B.start1 + B.length)) + (blk[b + 1].start2 - (B.start2 + B.length));
            al += g; gp += g;
        }
        for (int i = lane; i < B.length; i += 32) {
            const char c1 = rc1 ? complement_d(q1.at(q1.n - (B.start1 + i))) : q1.at(B.start1 - 1 + i);
            const char c2 = q2.at(B.start2 - 1 + i);
            ++al;
            id += same_nuc_d(c1, c2) ? 1 : 0;
            gp += (is_gap_d(c1) || is_gap_d(c2)) ? 1 : 0;
        }
    }
    n_ali = warp_sum(al); n_id = warp_sum(id); n_gap = warp_sum(gp);
}

// calcIdentity (AGEaligner.cpp:429-447): how far q keeps agreeing with itself shifted by bp2 - bp1, below bp1 and from bp1 on
__device__ int at_identity_d(const SeqD &q, int bp1, int bp2, int lane, int &left, int &right)
{
    left = right = 0;
    if (bp1 >= bp2) return 0;
    const int shift = bp2 - bp1;
    int down = 0, up = 0;
    for (int i0 = bp1 - 1;; i0 -= 32) {
        const int i = i0 - lane;
        const unsigned bad = __ballot_sync(0xFFFFFFFFu, !(i >= 0 && same_nuc_d(q.at(i), q.at(i + shift))));
        if (bad) { down += __ffs(bad) - 1; break; }
        down += 32;
    }
    for (int i0 = bp1;; i0 += 32) {
        const int i = i0 + lane;
        const unsigned bad = __ballot_sync(0xFFFFFFFFu, !(i < q.n && same_nuc_d(q.at(i), q.at(i + shift))));
        if (bad) { up += __ffs(bad) - 1; break; }
        up += 32;
    }
    left = -down; right = up;
    return up + down;
}

// Longest m <= n such that the m bases of q from `head` on equal the m bases that end at `tail` (the common core of
// calcOutsideIdentity / calcInsideIdentity, AGEaligner.cpp:392-427).  32 candidate lengths are checked side by side, the
// longest first; a lane gives up at its first mismatch.
__device__ int head_tail_overlap_d(const SeqD &q, int head, int tail, int n, int lane)
{
    for (int base = n; base >= 1; base -= 32) {
        const int m = base - lane;
        bool ok = m >= 1;
        for (int k = 0; __any_sync(0xFFFFFFFFu, ok && k < m); ++k)
            if (ok && k < m) ok = same_nuc_d(q.at(head + k), q.at(tail - m + 1 + k));
        const unsigned full = __ballot_sync(0xFFFFFFFFu, ok);
        if (full) return base - (__ffs(full) - 1);
    }
    return 0;
}
__device__ int outside_identity_d(const SeqD &q, int bp1, int bp2, int lane)
{
    return bp1 >= bp2 ? 0 : head_tail_overlap_d(q, bp2, bp1, min(bp1 + 1, q.n - bp2), lane);
}
__device__ int inside_identity_d(const SeqD &q, int bp1, int bp2, int lane)
{
    return bp1 >= bp2 ? 0 : head_tail_overlap_d(q, bp1, bp2, (bp2 - bp1 + 1) / 2, lane);
}

// AlignerOut::call of the alignment whose blocks were just written (left: nl blocks, right: nr blocks)
template <bool WIDE>
__device__ void alignment_numbers(const View<WIDE> &V, int flag, const Block *left, int nl, const Block *right, int nr, AlignerOut &O, int lane)
{
    const SeqD q1{V.T->raw1[V.half], V.len1}, q2{V.T->raw2[V.half], V.len2};
    int call[CALL_N];
#pragma unroll
    for (int i = 0; i < CALL_N; ++i) call[i] = 0;
    __syncwarp();                                           // the block lists were written by lane 0
    if (nl > 0) fragment_counts(left, nl, q1, q2, (flag & F_INVR) != 0, lane, call[CALL_ALIGNED], call[CALL_IDENTIC], call[CALL_GAP]);
    if (nr > 0) fragment_counts(right, nr, q1, q2, (flag & F_INVL) != 0, lane, call[CALL_ALIGNED + 1], call[CALL_IDENTIC + 1], call[CALL_GAP + 1]);
    if (nl > 0 && nr > 0) {
        // 0-based positions in the sequences as passed: a position p of the reverse complement is position len1 - p
        const Block le = left[nl - 1], rs = right[0];
        const int e1 = le.start1 + le.length - 1, e2 = le.start2 + le.length - 1;
        const int A = (flag & F_INVR) ? V.len1 - e1 : e1 - 1, B = (flag & F_INVL) ? V.len1 - rs.start1 : rs.start1 - 1;
        int s, e;                                           // the excised region (getExcisedRange, AGEaligner.cpp:156-177)
        if (flag & F_INDEL)      { s = A + 1; e = B - 1; }
        else if (flag & F_TDUP)  { s = B;     e = A; }
        else if (flag & F_INVL)  { s = A + 1; e = B; }
        else                     { s = A;     e = B - 1; }
        call[CALL_AT1] = at_identity_d(q1, s, e + 1, lane, call[CALL_AT1 + 1], call[CALL_AT1 + 2]);
        call[CALL_OUT1] = outside_identity_d(q1, s - 1, e + 1, lane);
        call[CALL_IN1] = inside_identity_d(q1, s, e, lane);
        const int L2 = e2 - 1, R2 = rs.start2 - 1;
        call[CALL_AT2] = at_identity_d(q2, L2 + 1, R2, lane, call[CALL_AT2 + 1], call[CALL_AT2 + 2]);
        call[CALL_OUT2] = outside_identity_d(q2, L2, R2, lane);
        call[CALL_IN2] = inside_identity_d(q2, L2 + 1, R2 - 1, lane);
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < CALL_N; ++i) O.call[i] = call[i];
    }
}

// K10-K12: the tracebacks of bpoint 0 and of the alternative printAlignment shows (:287-306) into the block pool.  A traceback
// is a latency-bound walk: the counting walk leaves its blocks in shared memory (`stage`: 2 x BLK_STAGE blocks of this warp) and
// they are copied into the pool once their place is known; only alignments with more blocks than that are walked twice.
constexpr int BLK_STAGE = 64;
template <bool WIDE>
__device__ void emit_blocks(const View<WIDE> &V, int flag, const BpList &L, AlignerOut &O, Block *pool, int pool_cap, int32_t *pool_used, int lane, Block *stage)
{
    int alt = -1;
    P2_PROF_T0(pc);
    Block *sl = stage, *sr = stage + BLK_STAGE;
    for (int which = 0; which < 2; ++which) {
        int idx = 0, nl = 0, nr = 0;
        if (which == 0) {
            nl = left_blocks(V, L.bp[0][0], L.bp[0][1], nullptr, 0, lane, sl, BLK_STAGE);
            nr = right_blocks(V, L.bp[0][2], L.bp[0][3], nullptr, 0, lane, sr, BLK_STAGE);
        } else {
            for (idx = L.n - 1; idx >= 1; --idx) {
                nl = left_blocks(V, L.bp[idx][0], L.bp[idx][1], nullptr, 0, lane, sl, BLK_STAGE);
                nr = right_blocks(V, L.bp[idx][2], L.bp[idx][3], nullptr, 0, lane, sr, BLK_STAGE);
                if (nl + nr > 0) break;
            }
            if (idx < 1) { nl = nr = 0; idx = -1; }
            alt = idx;
        }
        int off = 0;
        if (nl + nr > 0) {
            if (lane == 0) off = atomicAdd(pool_used, nl + nr);
            off = __shfl_sync(0xFFFFFFFFu, off, 0);
            if (off + nl + nr <= pool_cap) {
                __syncwarp();                                                     // the staged blocks were written by lane 0
                if (nl <= BLK_STAGE) { for (int i = lane; i < nl; i += 32) pool[off + nl - 1 - i] = sl[i]; }
                else left_blocks(V, L.bp[idx][0], L.bp[idx][1], pool + off, nl, lane);
                if (nr <= BLK_STAGE) { for (int i = lane; i < nr; i += 32) pool[off + nl + i] = sr[i]; }
                else right_blocks(V, L.bp[idx][2], L.bp[idx][3], pool + off + nl, nr, lane);
                __syncwarp();
                if (which == 0) alignment_numbers(V, flag, pool + off, nl, pool + off + nl, nr, O, lane);
            } else if (lane == 0) O.status = -4;          // pool overflow: host retries with a larger pool
        } else if (which == 0 && lane < CALL_N) O.call[lane] = 0;
        if (lane == 0) {
            O.n_blk[2 * which] = nl; O.n_blk[2 * which + 1] = nr;
            O.blk_off[2 * which] = off; O.blk_off[2 * which + 1] = off + nl;
        }
    }
    if (lane == 0) O.alt_index = alt;
    P2_PROF_ADD(3, pc);
    __syncwarp();
}

} // namespace age
