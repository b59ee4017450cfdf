// age_kernels.cu -- hand-written sm_100a kernels of the AGE realigner.
//
// Reference semantics (all citations: src/AsmvarAGE/src/AGEaligner.cpp of bioinformatics-centre/AsmVar):
//   K1/K2  calcScores   :978-1083   forward / reverse local fill, single matrix, trace-dependent gap cost
//   K3/K4  calcMaxima   :862-976    running prefix / suffix maxima and their trace
//   K6     findIndelBPs :540-613    K7 findInversionTDuplicationBPs :472-538
//   K8     traceBackMaxima :615-642 K9 addBpoints :644-668
//   K11/12 findLeftBlocks / findRightBlocks :752-860
//
// age_sweep_kernel   the hot kernel: per pair, the forward sweep streams the shifted running maxima D to HBM,
//                    the reverse sweep folds D back in and leaves one split maximum per (lane, window) tile.
//                    Neither sweep keeps a trace; both leave a 5 KB lane-state checkpoint per window.
// age_enum_kernel /  per pair, one warp: split maximum, -both / -inv winner selection, then for the winner only
//                    breakpoint enumeration and traceback.  Every trace bit it looks at comes from re-running
//                    the sweep over the one window that holds it (MODE_TRACE), from that window's checkpoint.
#include "age_sweep.cuh"
#include <cuda_runtime.h>
#include <cstdlib>
#include <cstdio>
#include <algorithm>

namespace age {

constexpr int SWEEP_WARPS = 4;
#ifndef SWEEP_OCC
#define SWEEP_OCC 4
#endif

// optional cycle accounting of pass 2 (AGE_P2_PROF=1): 0 maxima, 1 row ranges, 2 enumerate, 3 blocks, 4 forward re-sweeps,
// 5 reverse re-sweeps, 6 number of forward windows, 7 number of reverse windows
__device__ unsigned long long g_prof[8];
#define PROF_ADD(i, v) atomicAdd(&g_prof[i], (unsigned long long)(v))

// ------------------------------------------------------------------------------------------------
// pack: raw nucleotide strings -> the per-block column descriptors and per-row score tables of the sweep
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int nuc_code(char ch)            // Scorer.hh:24-39 with Sequence.hh:138-150,182-188
{
    switch (ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    case 'U': case 'u': return 4;
    default: return 5;                                       // scores 0 against everything
    }
}
__device__ __forceinline__ int comp_code(int c) { return c < 4 ? 3 - c : (c == 4 ? 0 : 5); }   // complement: U -> A

__global__ void __launch_bounds__(256)
age_pack_kernel(const PairTask *__restrict__ pairs, int n_pairs)
{
    for (int pi = blockIdx.x; pi < n_pairs; pi += gridDim.x) {
        const PairTask &T = pairs[pi];
        const int len1 = T.len1, len2 = T.len2, nb = T.nb;
        // column code of half k in direction dir: the forward fill of INVR and the reverse fill of INVL read the
        // reverse complement of seq1 (AGEaligner.cpp:1000, :1050)
        auto xcode = [&](int k, int dir, int c) -> int {
            const char *raw = T.raw1[k];
            if (!raw || c > len1) return 5;
            const bool rc = dir > 0 ? (T.flag[k] & F_INVR) : (T.flag[k] & F_INVL);
            return rc ? comp_code(nuc_code(raw[len1 - c])) : nuc_code(raw[c - 1]);
        };
        // .w of a descriptor = the N / IUPAC flags of the 32 blocks of its aligned group (bit b & 31 = block b has such a
        // column): the steady-state sweep takes its per-step branch from these bits and not from the descriptor it has just
        // loaded (age_sweep.cuh, sweep_window_steady)
        for (int dir = +1; dir >= -1; dir -= 2) {
            uint4 *dst = dir > 0 ? T.xblk_f : T.xblk_r;
            for (int base = 0; base < nb + 2; base += blockDim.x) {          // blockDim.x is a multiple of 32: warps cover aligned groups
                const int b = base + threadIdx.x;
                uint32_t sel[W] = {0, 0, 0, 0}, info = 0;
                if (b == 0 || b == nb + 1) info = 0x5B5B5B5Bu;
                else if (b <= nb) {
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        const int c = dir > 0 ? W * (b - 1) + 1 + w : W * (nb - b + 1) - w;   // columns beyond len1 are padding
                        const int a = xcode(0, dir, c), e = xcode(1, dir, c);
                        if (a < 4 && e < 4) sel[w] = (uint32_t)a | ((uint32_t)(a | 8) << 4) | ((uint32_t)(4 + e) << 8) | ((uint32_t)((4 + e) | 8) << 12);
                        else { sel[w] = 0; info |= (1u | ((uint32_t)a << 1) | ((uint32_t)e << 4)) << (8 * w); }
                    }
                }
                const uint32_t group = __ballot_sync(0xFFFFFFFFu, (info & 0x01010101u) != 0);
                if (b <= nb + 1) dst[b] = make_uint4(sel[0] | (sel[1] << 16), sel[2] | (sel[3] << 16), info, group);
            }
        }
        for (int q = threadIdx.x; q < T.P; q += blockDim.x) {
            uint32_t tab[2] = {0, 0};
            int code[2] = {5, 5};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (!T.raw2[k] || q >= len2) continue;
                const int yc = nuc_code(T.raw2[k][q]);
                code[k] = yc;
                if (yc >= 5) continue;
                for (int x = 0; x < 4; ++x) tab[k] |= ((uint32_t)(2 * (x == yc ? T.match[k] : T.mismatch[k])) & 0xFFu) << (8 * x);
            }
            T.ytab[q] = make_uint2(tab[0], tab[1]);
            T.ycode[q] = (uint8_t)(code[0] | (code[1] << 4));
        }
    }
}

void launch_pack(const PairTask *d_pairs, int n_pairs, cudaStream_t st)
{
    if (n_pairs <= 0) return;
    age_pack_kernel<<<std::min(n_pairs, 148 * 16), 256, 0, st>>>(d_pairs, n_pairs);
}

struct SweepSmem { WarpRings rings; DStage dstage; };

__global__ void __launch_bounds__(SWEEP_WARPS * 32, SWEEP_OCC)
age_sweep_kernel(const PairTask *__restrict__ pairs, int n_pairs, int32_t *counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SweepSmem *sm = reinterpret_cast<SweepSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= n_pairs) break;
        const PairTask &T = pairs[item];
        sweep_pair_fast<+1>(T, sm[w].rings, sm[w].dstage, lane);
        __threadfence_block();
        __syncwarp();
        sweep_pair_fast<-1>(T, sm[w].rings, sm[w].dstage, lane);
        cp_async_wait_all();
        __syncwarp();
    }
}

// Long alignments / small batches: the NW warps of a CTA share one pair and run its strips as a pipeline
// (warp w owns strips w, w + NW, ...; strip si trails strip si-1 by 32+ steps through the boundary rows in global memory).
constexpr int MAX_STRIPS_LONG = 128;

__global__ void __launch_bounds__(512, 1)
age_sweep_long_kernel(const PairTask *__restrict__ pairs, int n_pairs, int32_t *counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SweepSmem *sm = reinterpret_cast<SweepSmem *>(smem_raw);
    __shared__ int prog[MAX_STRIPS_LONG];
    __shared__ int s_item;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(counter, 1);
        for (int i = threadIdx.x; i < MAX_STRIPS_LONG; i += blockDim.x) prog[i] = 0;
        __syncthreads();
        const int item = s_item;
        if (item >= n_pairs) break;
        const PairTask &T = pairs[item];
        sweep_pair_fast<+1>(T, sm[w].rings, sm[w].dstage, lane, w, nw, prog);
        __threadfence();
        __syncthreads();
        for (int i = threadIdx.x; i < MAX_STRIPS_LONG; i += blockDim.x) prog[i] = 0;
        __syncthreads();
        sweep_pair_fast<-1>(T, sm[w].rings, sm[w].dstage, lane, w, nw, prog);
        cp_async_wait_all();
    }
}

void launch_sweeps(const PairTask *d_pairs, int n_pairs, int max_strips, int32_t *d_counter, cudaStream_t st)
{
    if (n_pairs <= 0) return;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaMemsetAsync(d_counter, 0, sizeof(int32_t), st);
    // one warp per pair fills the GPU from ~sms * 16 pairs on; below that, share each pair among up to 16 warps
    int nw = 1;
    const int force_nw = getenv("AGE_LONG_NW") ? atoi(getenv("AGE_LONG_NW")) : 0;   // test knob: 1 = one warp per pair, 2..16 = pipelined
    while (nw < 16 && nw * 2 <= max_strips && n_pairs * nw * 2 <= sms * 16) nw *= 2;
    if (force_nw >= 1 && force_nw <= 16 && (force_nw & (force_nw - 1)) == 0) nw = force_nw;
    if (max_strips > MAX_STRIPS_LONG) nw = 1;
    if (nw > 1) {
        const int smem = nw * (int)sizeof(SweepSmem);
        cudaFuncSetAttribute(age_sweep_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const int per_sm = std::max(1, std::min(16 / nw, 4));
        age_sweep_long_kernel<<<std::min(n_pairs, sms * per_sm), nw * 32, smem, st>>>(d_pairs, n_pairs, d_counter);
        return;
    }
    int blocks = (n_pairs + SWEEP_WARPS - 1) / SWEEP_WARPS;
    if (blocks > sms * SWEEP_OCC) blocks = sms * SWEEP_OCC;  // persistent: the work counter feeds the warps
    const int smem = SWEEP_WARPS * (int)sizeof(SweepSmem);
    cudaFuncSetAttribute(age_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    age_sweep_kernel<<<blocks, SWEEP_WARPS * 32, smem, st>>>(d_pairs, n_pairs, d_counter);
}

// ------------------------------------------------------------------------------------------------
// pass 2
// ------------------------------------------------------------------------------------------------
struct View {
    const PairTask *T;
    P2Scratch sc;
    WarpRings *R;
    TraceSmem *X;
    int len1, len2, nb, nsteps, nwin, nstrips, half, lane, max2, vwords;

    // ---- geometry --------------------------------------------------------------------------------
    __device__ int fstep(int q, int c) const { return (c - 1) / W + 1 + ((q & 255) >> 3); }          // forward step of cell (q+1, c)
    __device__ int rstep(int q, int c) const { return nb - (c - 1) / W + 31 - ((q & 255) >> 3); }    // reverse step
    __device__ size_t slot(int q, int t) const { return ((size_t)(q >> 8) * nsteps + (t - 1)) * 32 + ((q & 255) >> 3); }
    __device__ bool interior(int r, int c) const { return r >= 1 && c >= 1 && r <= len2 && c <= len1; }

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
        const long long c0 = clock64();
        const TraceOut TO{sc.P[plane], sc.hot, with_mp ? sc.Mp : nullptr, 16 * half, max2};
        switch (plane) {
        case 0:  sweep_window_trace<+1, MODE_TRACE_S>(*T, *R, *X, lane, s, win, TO); break;
        case 1:  sweep_window_trace<+1, MODE_TRACE_M>(*T, *R, *X, lane, s, win, TO); break;
        case 2:  sweep_window_trace<-1, MODE_TRACE_S>(*T, *R, *X, lane, s, win, TO); break;
        default: sweep_window_trace<-1, MODE_TRACE_M>(*T, *R, *X, lane, s, win, TO); break;
        }
        set_valid(plane, s, win);
        if (with_mp) set_valid(4, s, win);
        if (lane == 0) PROF_ADD(4 + (plane >> 1), clock64() - c0);
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
        if (T->dfmt == DFMT_RAW) return T->D[(((slot * W + w) * 2 + (k >> 2)) * 32 + l) * 4 + (k & 3)];
        const uint32_t *q = T->D + (slot * 5 * 32 + l) * 4;                      // OFF8: entry 0 of the column + the row's offset bytes
        const uint32_t wd = q[(size_t)(1 + w) * 128 + (k >> 1)] >> (8 * (k & 1));       // bytes {A_2j, A_2j+1, B_2j, B_2j+1}
        return q[w] + (wd & 0x00FF00FFu);
    }
    __device__ int F2(int r, int c) const { return (F2w(r, c) >> (16 * half)) & 0xFFFF; }
    __device__ int R2(int r, int c) const {              // 2 * Rmax[r][c]; needs ensure_rm(strip of r)
        if (r < 1 || c < 1 || r > len2 || c > len1) return 0;
        if (c == 1) return (T->Rfirst[r] >> (16 * half)) & 0xFFFF;
        const int q = r - 1, l = (q & 255) >> 3, k = q & 7, t = rstep(q, c), w = W * ((c - 1) / W + 1) - c;
        return (sc.Mp[(((((size_t)(q >> 8) * nsteps + (t - 1)) * W + w) * 2 + (k >> 2)) * 32 + l) * 4 + (k & 3)] >> (16 * half)) & 0xFFFF;
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
        const uint2 v = sc.P[plane][slot(q, rev ? rstep(q, c) : fstep(q, c))];
        const int w = rev ? W * ((c - 1) / W + 1) - c : (c - 1) % W;
        return (((w < 2 ? v.x : v.y) >> (16 * (w & 1))) >> (2 * (q & 7))) & 3u;
    }
    __device__ bool hot_nc(int r, int c) const {          // reverse cell (r, c): Fmax[r-1][c-1] + Rmax[r][c] == max
        const int q = r - 1;
        const int w = W * ((c - 1) / W + 1) - c;
        return (sc.hot[slot(q, rstep(q, c))] >> (8 * w + (q & 7))) & 1u;
    }
    // Starting at (r, c) and stepping by (dr, dc): how many consecutive cells (at most 32) carry code t, and the
    // code of the first cell that does not (meaningful when the result is < 32).
    __device__ int run_len(int plane, int r, int c, int dr, int dc, uint32_t t, uint32_t &tnext) const {
        const int rr = r + lane * dr, cc = c + lane * dc;
        ensure_lanes(plane, rr, cc);
        const uint32_t code = code_nc(plane, rr, cc);
        const unsigned m = __ballot_sync(0xFFFFFFFFu, code != t);
        const int n = m ? __ffs(m) - 1 : 32;
        tnext = __shfl_sync(0xFFFFFFFFu, code, n & 31);
        return n;
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
        for (int depth = 0; depth < 64; ++depth) {
            ensure_lanes(plane, run0 ? r + depth * dir : -1, x);
            ensure_lanes(plane, run1 ? r + depth * dir : -1, x + dir);
            if (run0) { c0 = code_nc(plane, r + depth * dir, x); if (c0 == T_VERT) d0 = depth + 1; else run0 = false; }
            if (run1) { c1 = code_nc(plane, r + depth * dir, x + dir); if (c1 == T_VERT) d1 = depth + 1; else run1 = false; }
            if (!__any_sync(0xFFFFFFFFu, run0 || run1)) break;
        }
        if (!in || run0 || run1) return false;            // deeper ladders than 64 rows are walked one by one
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
__device__ void walk_chain(const View &V, int plane, int sign, int &c, int &r)
{
    uint32_t t = V.code_at(plane, r, c);
    while (t != T_STOP) {
        const int dr = t == T_VERT ? sign : 0, dc = t == T_VERT ? 0 : sign;
        uint32_t tn;
        const int n = V.run_len(plane, r, c, dr, dc, t, tn);
        r += n * dr; c += n * dc;
        t = n < 32 ? tn : V.code_at(plane, r, c);
    }
}
__device__ void walk_left(const View &V, int &lg1, int &lg2) { walk_chain(V, 1, -1, lg1, lg2); }
__device__ void walk_right(const View &V, int &rg1, int &rg2) { walk_chain(V, 3, +1, rg1, rg2); }

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
// Fmax + Rmax == max (from the hot windows of the reverse sweep and the border rule); rows with rhi < rlo are skipped.
__device__ void indel_enumerate(const View &V, BpList &L, const int *rlo, const int *rhi, int lane)
{
    const int len2 = V.len2;
    int n_add = 0;
    bool stop = false;
    // Most rows hold no candidate: the per-row column ranges are inspected 32 rows at a time (one per lane) and only the
    // non-empty rows are visited, in the reference's order.
    for (int rb = 0; rb <= len2 && !stop; rb += 32) {                     // forward scan (:552-573)
      const int myr = min(rb + lane, len2), mylo = rlo[myr], myhi = rhi[myr];
      const bool mine = rb + lane <= len2 && myhi >= mylo;
      unsigned rows = __ballot_sync(0xFFFFFFFFu, mine);
      if (rows && __all_sync(0xFFFFFFFFu, !mine || myhi - mylo < 4)) {
        // at most four candidate columns per row (the split runs along a column: insertions): 8 rows x 4 columns are tested
        // side by side, lane order = row-major order, hits are taken in that order
        for (int sub = 0; sub < 32 && !stop; sub += 8) {
            if (!((rows >> sub) & 0xFFu)) continue;
            const int src = sub + (lane >> 2);
            const int r = rb + src, lo = __shfl_sync(0xFFFFFFFFu, mylo, src), hi = __shfl_sync(0xFFFFFFFFu, myhi, src);
            const int c = lo + (lane & 3);
            const bool in = ((rows >> src) & 1u) && c <= hi;
            V.ensure_lanes(3, in ? r + 1 : -1, c + 1);
            const bool tie = in && V.tie_nc(r, c);
            V.ensure_lanes(1, tie ? r : -1, c);
            unsigned m = __ballot_sync(0xFFFFFFFFu, tie && V.code_nc(1, r, c) == T_STOP);
            while (m && !stop) {
                const int b = __ffs(m) - 1; m &= m - 1;
                const int cc = __shfl_sync(0xFFFFFFFFu, c, b), rr = rb + sub + (b >> 2);
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
        for (int cb = lo & ~31; cb <= hi && !stop; cb += 32) {
            const int c = cb + lane;
            const bool in = c >= lo && c <= hi;
            V.ensure_lanes(3, in ? r + 1 : -1, c + 1);
            const bool tie = in && V.tie_nc(r, c);
            V.ensure_lanes(1, tie ? r : -1, c);
            const bool hit = tie && V.code_nc(1, r, c) == T_STOP;
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
    }
    n_add = 0; stop = false;
    for (int rb = len2 & ~31; rb >= 0 && !stop; rb -= 32) {               // reverse scan (:586-610)
      const int myr = min(rb + lane, len2), mylo = rlo[myr], myhi = rhi[myr];
      const bool mine = rb + lane <= len2 && myhi >= mylo;
      unsigned rows = __ballot_sync(0xFFFFFFFFu, mine);
      if (rows && __all_sync(0xFFFFFFFFu, !mine || myhi - mylo < 4)) {
        for (int sub = 24; sub >= 0 && !stop; sub -= 8) {
            if (!((rows >> sub) & 0xFFu)) continue;
            const int src = sub + (lane >> 2);
            const int r = rb + src, lo = __shfl_sync(0xFFFFFFFFu, mylo, src), hi = __shfl_sync(0xFFFFFFFFu, myhi, src);
            const int c = lo + (lane & 3);
            const bool in = ((rows >> src) & 1u) && c <= hi;
            V.ensure_lanes(3, in ? r + 1 : -1, c + 1);
            unsigned m = __ballot_sync(0xFFFFFFFFu, in && V.tie_nc(r, c) && V.code_nc(3, r + 1, c + 1) == T_STOP);
            while (m && !stop) {
                const int b = 31 - __clz(m); m &= ~(1u << b);
                const int cc = __shfl_sync(0xFFFFFFFFu, c, b), rr = rb + sub + (b >> 2);
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
        for (int cb = hi & ~31; cb >= (lo & ~31) && !stop; cb -= 32) {
            const int c = cb + lane;
            const bool in = c >= lo && c <= hi;
            V.ensure_lanes(3, in ? r + 1 : -1, c + 1);
            const bool hit = in && V.tie_nc(r, c) && V.code_nc(3, r + 1, c + 1) == T_STOP;
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
}

// Hot-window refinement: every reverse window that holds a (lane, window) tile whose maximum equals the split
// maximum is re-swept with the trace on; its hot bits widen the column range of the K6 rows they belong to.
__device__ void indel_row_ranges(const View &V, int *rlo, int *rhi, int lane)
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
    for (int hi = 0; hi < nhot; ++hi) {
        const int wi = V.sc.hotlist[1 + hi];
        const int s = wi / T.nwin, win = wi % T.nwin;
        V.ensure(3, s, win);
        const int t0 = CKS * win + 1, t1 = min(CKS * win + CKS, T.nsteps);
        uint32_t hw[CKS];
#pragma unroll
        for (int i = 0; i < CKS; ++i) {                               // the lane's hot words of all steps of the window, in flight together
            const int t = t0 + i, bq = t - 31 + lane;                 // logical reverse block of this lane in step t
            hw[i] = (t <= t1 && bq >= 1 && bq <= nb) ? __ldcg(V.sc.hot + ((size_t)s * T.nsteps + (t - 1)) * 32 + lane) : 0u;
        }
        // the hot cells of a lane lie in its own 8 rows: column ranges are gathered in registers, no atomics
        int lo8[RPT], hi8[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) { lo8[k] = 0x7FFFFFFF; hi8[k] = -1; }
#pragma unroll
        for (int i = 0; i < CKS; ++i) {
            const uint32_t x = hw[i];
            if (x == 0) continue;
            const int cb = W * (nb + 1 - (t0 + i - 31 + lane));       // reverse cell column of block column w is cb - w
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const int c = cb - w;
                if (c > len1) continue;                               // padding
                const uint32_t byte = (x >> (8 * w)) & 0xFFu;
#pragma unroll
                for (int k = 0; k < RPT; ++k)
                    if ((byte >> k) & 1u) { lo8[k] = min(lo8[k], c - 1); hi8[k] = max(hi8[k], c - 1); }
            }
        }
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = s * STRIP + lane * RPT + k + 1;             // reverse cell row; K6 row r - 1
            if (hi8[k] >= 0 && r <= len2) { rlo[r - 1] = min(rlo[r - 1], lo8[k]); rhi[r - 1] = max(rhi[r - 1], hi8[k]); }
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

// K7 findInversionTDuplicationBPs (:472-538).  The reference walks every (j1, j2) pair; here the monotone
// rows are binary-searched for the run of matching values, and of the start cells whose traceBackMaxima walks
// merge (View::merges_into_next: they share the endpoint, hence the addBpoints key) only the last one is walked --
// the others are guaranteed duplicates (:652-655) and change nothing.
__device__ void invtdup_enumerate(const View &V, BpList &L, int lane)
{
    const int len1 = V.len1, len2 = V.len2, max2 = V.max2;
    const uint32_t *colF = V.T->Dlast, *colR = V.T->Rfirst;
    const int sh = 16 * V.half;
    auto A = [&](int r) { return r >= 1 ? (int)((colF[r] >> sh) & 0xFFFF) : 0; };                 // 2*Fmax[r][len1]
    auto B = [&](int r) { return (r >= 1 && r <= len2) ? (int)((colR[r] >> sh) & 0xFFFF) : 0; };  // 2*Rmax[r][1]
    int n_add = 0;
    bool stop = false;
    long long kc = clock64();
    for (int r = 0; r <= len2 && !stop; ++r) {                            // forward (:485-504)
        if (A(r) + B(r + 1) != max2) continue;
        if (r + 1 <= len2) V.ensure_rm(r >> 8);                           // row r+1 of Rmax and RM
        for (int jb = 0; jb <= len1 && !stop; jb += 32) {
            const int j = jb + lane;
            V.ensure_lanes(1, j <= len1 ? r : -1, j);
            unsigned m = __ballot_sync(0xFFFFFFFFu, j <= len1 && V.code_nc(1, r, j) == T_STOP);
            while (m && !stop) {
                const int j1 = jb + __ffs(m) - 1; m &= m - 1;
                const int tv = max2 - V.F2(r, j1);
                if (tv < 0) continue;
                // columns x in [1, len1+1] of row r+1, values non-increasing: first x with value <= tv
                int lo = 1, hi = len1 + 2;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (V.R2(r + 1, mid) <= tv) hi = mid; else lo = mid + 1; }
                if (V.R2(r + 1, lo) != tv) continue;
                // the run of columns with value tv ends before the first x with value < tv
                int l2 = lo, h2 = len1 + 2;
                while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (V.R2(r + 1, mid) < tv) h2 = mid; else l2 = mid + 1; }
                const int xhi = l2 - 1;
                for (int xb = lo; xb <= xhi && !stop; xb += 32) {
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
        }
    }
    if (lane == 0) PROF_ADD(1, clock64() - kc);
    kc = clock64();
    n_add = 0; stop = false;
    for (int i2 = len2 + 1; i2 >= 1 && !stop; --i2) {                     // reverse (:515-534)
        if (A(i2 - 1) + B(i2) != max2) continue;
        if (i2 <= len2) V.ensure_rm((i2 - 1) >> 8);
        for (int jb = ((len1 + 1) / 32) * 32; jb >= 0 && !stop; jb -= 32) {
            const int j = jb + lane;
            V.ensure_lanes(3, (j >= 1 && j <= len1 + 1) ? i2 : -1, j);
            unsigned m = __ballot_sync(0xFFFFFFFFu, j >= 1 && j <= len1 + 1 && V.code_nc(3, i2, j) == T_STOP);
            while (m && !stop) {
                const int b = 31 - __clz(m); m &= ~(1u << b);
                const int j2 = jb + b;
                const int tv = max2 - V.R2(i2, j2);
                if (tv < 0) continue;
                // columns x in [0, len1] of row i2-1, values non-decreasing: last x with value <= tv
                int lo = -1, hi = len1;
                while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (V.F2(i2 - 1, mid) <= tv) lo = mid; else hi = mid - 1; }
                if (lo < 0 || V.F2(i2 - 1, lo) != tv) continue;
                // the run of columns with value tv starts after the last x with value < tv
                int l2 = -1, h2 = lo;
                while (l2 < h2) { const int mid = (l2 + h2 + 1) >> 1; if (V.F2(i2 - 1, mid) < tv) l2 = mid; else h2 = mid - 1; }
                const int xlo = l2 + 1;
                for (int xb = lo; xb >= xlo && !stop; xb -= 32) {
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
        }
    }
    if (lane == 0) PROF_ADD(2, clock64() - kc);
}

// K11 findLeftBlocks (:752-808).  dst == nullptr: count only.  Blocks are produced last-first; a block is a maximal
// run of DIAGONAL codes (any H / V step in between starts a new block, :774-790).
__device__ int left_blocks(const View &V, int c, int r, Block *dst, int n_total, int lane)
{
    int n = 0, run = 0;
    uint32_t t = V.code_at(0, r, c);
    while (t != T_STOP) {
        const int dr = t == T_HORI ? 0 : -1, dc = t == T_VERT ? 0 : -1;
        uint32_t tn;
        const int len = V.run_len(0, r, c, dr, dc, t, tn);
        r += len * dr; c += len * dc;
        if (t == T_DIAG) run += len;
        else if (run > 0) { if (dst && lane == 0) dst[n_total - 1 - n] = Block{c - len * dc + 1, r - len * dr + 1, run}; ++n; run = 0; }
        t = len < 32 ? tn : V.code_at(0, r, c);
    }
    if (run > 0) { if (dst && lane == 0) dst[n_total - 1 - n] = Block{c + 1, r + 1, run}; ++n; }
    return n;
}

// K12 findRightBlocks (:810-860)
__device__ int right_blocks(const View &V, int c, int r, Block *dst, int lane)
{
    int n = 0, run = 0, s1 = 0, s2 = 0;
    uint32_t t = V.code_at(2, r, c);
    while (t != T_STOP) {
        const int dr = t == T_HORI ? 0 : 1, dc = t == T_VERT ? 0 : 1;
        uint32_t tn;
        const int len = V.run_len(2, r, c, dr, dc, t, tn);
        if (t == T_DIAG) { if (run == 0) { s1 = c; s2 = r; } run += len; }     // interior cells only (:820 always holds)
        else if (run > 0) { if (dst && lane == 0) dst[n] = Block{s1, s2, run}; ++n; run = 0; }
        r += len * dr; c += len * dc;
        t = len < 32 ? tn : V.code_at(2, r, c);
    }
    if (run > 0) { if (dst && lane == 0) dst[n] = Block{s1, s2, run}; ++n; }
    return n;
}

constexpr int P2_WARPS = 4;
#ifndef P2_OCC
#define P2_OCC 4
#endif

__device__ P2Scratch carve_set(const PairTask &T, int set)
{
    const size_t slots = (size_t)T.nstrips * T.nsteps, nwt = (size_t)T.nstrips * T.nwin;
    char *base = T.p2set + (size_t)set * T.p2stride;
    P2Scratch sc;
    for (int i = 0; i < 4; ++i) { sc.P[i] = (uint2 *)base; base += p2_up(slots * 32 * 8); }
    sc.hot = (uint32_t *)base; base += p2_up(slots * 32 * 4);
    sc.Mp = T.k7 ? (uint32_t *)base : nullptr; base += T.k7 ? p2_up(slots * W * 2 * 32 * 16) : 0;
    sc.vwords = (int)p2_vwords(nwt);
    sc.valid = (uint32_t *)base; base += p2_up((size_t)9 * sc.vwords * 4);
    sc.rowrange = (int32_t *)base; base += p2_up(2 * ((size_t)T.len2 + 2) * 4);
    sc.hotlist = (int32_t *)base;
    return sc;
}

// ---- select: split maxima, -both / -inv winner, work list of the windows along the split row / column -------------
constexpr int SEL_WARPS = 4;

__global__ void __launch_bounds__(SEL_WARPS * 32)
age_select_kernel(Pass2Params prm)
{
    const int lane = threadIdx.x & 31;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(prm.next + 0, 1);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= prm.n_pairs) break;
        const PairTask &T = prm.pairs[item];
        const int len2 = T.len2;
        // split maxima of both halves (:545-550 / :476-481)
        uint32_t mx_indel = 0, mx_col = 0;                       // packed u16x2: max of 2*(Fmax + Rmax)
        const int nwt = T.nstrips * T.nwin;
        for (int i = lane; i < nwt * 32; i += 32) mx_indel = __vmaxu2(mx_indel, T.tilemax[i]);
        for (int r = lane; r <= len2; r += 32) {
            const uint32_t f = r >= 1 ? T.Dlast[r] : 0u, q = r + 1 <= len2 ? T.Rfirst[r + 1] : 0u;
            mx_col = __vmaxu2(mx_col, __vadd2(f, q));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            mx_indel = __vmaxu2(mx_indel, __shfl_xor_sync(0xFFFFFFFFu, mx_indel, o));
            mx_col   = __vmaxu2(mx_col,   __shfl_xor_sync(0xFFFFFFFFu, mx_col, o));
        }
        mx_indel = __vmaxu2(mx_indel, T.Dlast[len2]);            // Fmax[len2][len1] + 0 (border cells)
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
        if (lane == 0) prm.hdr[item] = PairHdr{{max2[0], max2[1]}, {need[0] ? 1 : 0, need[1] ? 1 : 0}};
        for (int h = 0; h < 2; ++h) {
            if (!need[h]) continue;
            const P2Scratch sc = carve_set(T, T.select ? 0 : h);
            for (int i = lane; i < 9 * sc.vwords; i += 32) sc.valid[i] = 0;
            __syncwarp();
            if (T.flag[h] != F_INDEL) {
                // K7: the rows that tie at the maximum are scanned along their whole length (:485-534): queue every
                // window of their strips -- the forward maxima trace of row r, the reverse one (with the maxima
                // themselves) of row r+1
                uint32_t *qf = sc.valid + 5 * sc.vwords, *qr = sc.valid + 6 * sc.vwords;
                const int sh = 16 * h;
                for (int rb = 0; rb <= len2; rb += 32) {
                    const int r = rb + lane;
                    bool tie = false;
                    if (r <= len2) {
                        const int a = r >= 1 ? (int)((T.Dlast[r] >> sh) & 0xFFFF) : 0, b = r + 1 <= len2 ? (int)((T.Rfirst[r + 1] >> sh) & 0xFFFF) : 0;
                        tie = a + b == max2[h];
                    }
                    unsigned m = __ballot_sync(0xFFFFFFFFu, tie);
                    while (m) {
                        const int rr = rb + __ffs(m) - 1; m &= m - 1;
                        for (int pass = 0; pass < 2; ++pass) {
                            const int row = pass == 0 ? rr : rr + 1;                 // 1-based row whose strip is needed
                            if (row < 1 || row > len2) continue;
                            const int s = (row - 1) >> 8;
                            uint32_t *qd = pass == 0 ? qf : qr;
                            for (int w2 = lane; w2 < T.nwin; w2 += 32) {
                                const int i = s * T.nwin + w2;
                                if (atomicOr(&qd[i >> 5], 1u << (i & 31)) & (1u << (i & 31))) continue;
                                const int at = atomicAdd(prm.queue_n, 1);
                                if (at < prm.queue_cap) prm.queue[at] = make_uint2((uint32_t)item, (uint32_t)(h | ((pass == 0 ? 1 : 5) << 1) | (s << 4) | (w2 << 16)));
                            }
                        }
                    }
                }
                continue;
            }
            if (max2[h] == 0) continue;
            // every reverse window with a tile at the maximum, and the forward windows over the same cells
            uint32_t *queued = sc.valid + 5 * sc.vwords;
            int nhot = 0;
            for (int wi = 0; wi < nwt; ++wi) {
                const bool hot = (int)((T.tilemax[(size_t)wi * 32 + lane] >> (16 * h)) & 0xFFFF) == max2[h];
                if (!__any_sync(0xFFFFFFFFu, hot)) continue;
                if (lane == 0) sc.hotlist[1 + nhot] = wi;
                ++nhot;
                if (lane >= 3) continue;
                const int s = wi / T.nwin, win = wi % T.nwin;
                // reverse steps t' of the window meet the forward steps nb+32-t' (same anti-diagonal)
                const int t1 = min(CKS * win + CKS, T.nsteps);
                const int fa = min(max((T.nb + 32 - t1 - 1) / CKS, 0), T.nwin - 1), fb = min(max((T.nb + 31 - CKS * win - 1) / CKS, 0), T.nwin - 1);
                const int plane = lane == 0 ? 3 : 1, w2 = lane == 0 ? win : (lane == 1 ? fa : fb);
                if (lane == 2 && fb == fa) continue;
                if (plane == 1) {                                 // forward windows repeat between neighbours: queue once
                    const int i = s * T.nwin + w2;
                    if (atomicOr(&queued[i >> 5], 1u << (i & 31)) & (1u << (i & 31))) continue;
                }
                const int at = atomicAdd(prm.queue_n, 1);
                if (at < prm.queue_cap) prm.queue[at] = make_uint2((uint32_t)item, (uint32_t)(h | (plane << 1) | (s << 4) | (w2 << 16)));
            }
            if (lane == 0) sc.hotlist[0] = nhot;
        }
    }
}

// ---- pre-sweep: one warp per queued window, at full occupancy ------------------------------------------------------
constexpr int PRE_WARPS = 4;
#ifndef PRE_OCC
#define PRE_OCC 4
#endif

__global__ void __launch_bounds__(PRE_WARPS * 32, PRE_OCC)
age_presweep_kernel(Pass2Params prm, int round)
{
    __shared__ WarpRings rings[PRE_WARPS];
    __shared__ TraceSmem tsm[PRE_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = min(*prm.queue_n, prm.queue_cap);
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(prm.next + (round ? 4 : 1), 1);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= n) break;
        const uint2 q = prm.queue[item];
        const PairTask &T = prm.pairs[q.x];
        const int h = q.y & 1, plane = (q.y >> 1) & 7, s = (q.y >> 4) & 0xFFF, win = q.y >> 16;
        const P2Scratch sc = carve_set(T, T.select ? 0 : h);
        const bool with_mp = plane == 5;                     // reverse maxima trace + the maxima themselves (K7)
        const int pl = plane == 5 ? 3 : plane;
        const TraceOut TO{sc.P[pl], sc.hot, with_mp ? sc.Mp : nullptr, 16 * h, prm.hdr[q.x].max2[h]};
        switch (pl) {
        case 0:  sweep_window_trace<+1, MODE_TRACE_S>(T, rings[w], tsm[w], lane, s, win, TO); break;
        case 1:  sweep_window_trace<+1, MODE_TRACE_M>(T, rings[w], tsm[w], lane, s, win, TO); break;
        case 2:  sweep_window_trace<-1, MODE_TRACE_S>(T, rings[w], tsm[w], lane, s, win, TO); break;
        default: sweep_window_trace<-1, MODE_TRACE_M>(T, rings[w], tsm[w], lane, s, win, TO); break;
        }
        __threadfence();
        __syncwarp();
        const int i = s * T.nwin + win;
        if (lane == 0) {
            atomicOr(&sc.valid[pl * sc.vwords + (i >> 5)], 1u << (i & 31));
            if (with_mp) atomicOr(&sc.valid[4 * sc.vwords + (i >> 5)], 1u << (i & 31));
        }
    }
}

// Queue the windows a traceback from (r0, c0) will probably visit: the path runs along the diagonal (dir = -1: up-left
// through the forward scores trace, plane 0; dir = +1: down-right through the reverse one, plane 2) for about as many
// steps as the fragment scores.  A wrong guess costs nothing but time: windows the walk needs and that were not
// pre-swept are re-swept on demand.
__device__ void queue_path(const Pass2Params &prm, const View &V, int item, int plane, int r0, int c0, int dir, int steps)
{
    uint32_t *queued = V.sc.valid + (plane == 0 ? 7 : 8) * V.vwords;
    steps = min(steps, dir < 0 ? min(r0, c0) : min(V.len2 - r0, V.len1 - c0) + 1);
    for (int i0 = 0; i0 < steps; i0 += 32 * RPT) {
        const int i = i0 + V.lane * RPT;
        if (i >= steps) continue;
        const int r = r0 + dir * i;
        for (int dc = -12; dc <= 12; dc += 12) {               // gaps shift the path off the diagonal by a few columns
            const int c = c0 + dir * i + dc;
            if (!V.interior(r, c)) continue;
            const int s = (r - 1) >> 8, win = V.win_of(plane, r, c), bit = s * V.nwin + win;
            if (atomicOr(&queued[bit >> 5], 1u << (bit & 31)) & (1u << (bit & 31))) continue;
            const int at = atomicAdd(prm.queue_n, 1);
            if (at < prm.queue_cap) prm.queue[at] = make_uint2((uint32_t)item, (uint32_t)(V.half | (plane << 1) | (s << 4) | (win << 16)));
        }
    }
}

// ---- enumerate: breakpoints of the selected aligner(s) (K5-K9), then the work list of the traceback windows ----------
__global__ void __launch_bounds__(P2_WARPS * 32, P2_OCC)
age_enum_kernel(Pass2Params prm)
{
    __shared__ int bp_s[P2_WARPS][MAX_BP][4];
    __shared__ WarpRings rings[P2_WARPS];
    __shared__ TraceSmem tsm[P2_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(prm.next + 2, 1);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= prm.n_pairs) break;
        const PairTask &T = prm.pairs[item];
        const int len1 = T.len1, len2 = T.len2;
        const PairHdr hd = prm.hdr[item];
        long long pc = clock64();
        const long long pair_c0 = pc;
        for (int h = 0; h < 2; ++h) {
            if (!hd.need[h]) continue;
            const P2Scratch sc = carve_set(T, T.select ? 0 : h);
            const View V{&T, sc, &rings[w], &tsm[w], len1, len2, T.nb, T.nsteps, T.nwin, T.nstrips, h, lane, hd.max2[h], sc.vwords};
            AlignerOut &O = prm.outs[T.out[h]];
            BpList L{0, bp_s[w]};
            __syncwarp();
            if (T.flag[h] == F_INDEL) {
                int *rlo = sc.rowrange, *rhi = sc.rowrange + (len2 + 2);
                pc = clock64();
                indel_row_ranges(V, rlo, rhi, lane);
                if (lane == 0) { PROF_ADD(1, clock64() - pc); atomicMax(&g_prof[6], ((unsigned long long)((clock64() - pc) >> 10) << 32) | ((unsigned long long)min(sc.hotlist[0], 4095) << 16) | (unsigned long long)(len2 & 0xFFFF)); }
                pc = clock64();
                indel_enumerate(V, L, rlo, rhi, lane);
                if (lane == 0) { PROF_ADD(2, clock64() - pc); atomicMax(&g_prof[7], (unsigned long long)(clock64() - pc)); }
            } else invtdup_enumerate(V, L, lane);
            __syncwarp();
            if (lane == 0) {
                O.status = 0;
                O.score = hd.max2[h] >> 1;
                O.n_bp = L.n;
            }
            for (int i = lane; i < L.n * 4; i += 32) O.bp[i >> 2][i & 3] = L.bp[i >> 2][i & 3];
            // the tracebacks of bpoint 0 and of the alternative printAlignment shows (the last one, normally)
            for (int which = 0; which < 2; ++which) {
                const int idx = which == 0 ? 0 : L.n - 1;
                if (which == 1 && idx < 1) break;
                const int lg1 = L.bp[idx][0], lg2 = L.bp[idx][1], rg1 = L.bp[idx][2], rg2 = L.bp[idx][3];
                const int fl = V.F2(lg2, lg1) >> 1, fr = max((hd.max2[h] >> 1) - fl, 0);     // scores of the two fragments
                queue_path(prm, V, item, 0, lg2, lg1, -1, fl + fl / 4 + 64);
                queue_path(prm, V, item, 2, rg2, rg1, +1, fr + fr / 4 + 64);
            }
            __syncwarp();
        }
        if (lane == 0) atomicMax(&g_prof[0], ((unsigned long long)(clock64() - pair_c0) >> 10 << 24) | (unsigned long long)item);
    }
}

// ---- blocks: K10-K12 tracebacks of bpoint 0 and of the alternative shown by printAlignment (:287-306) ---------------------
// (the walks are executed by the whole warp: a trace window that is not resident is re-swept by all lanes)
__global__ void __launch_bounds__(P2_WARPS * 32, P2_OCC)
age_blocks_kernel(Pass2Params prm)
{
    __shared__ int bp_s[P2_WARPS][MAX_BP][4];
    __shared__ WarpRings rings[P2_WARPS];
    __shared__ TraceSmem tsm[P2_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(prm.next + 3, 1);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= prm.n_pairs) break;
        const PairTask &T = prm.pairs[item];
        const PairHdr hd = prm.hdr[item];
        for (int h = 0; h < 2; ++h) {
            if (!hd.need[h]) continue;
            const P2Scratch sc = carve_set(T, T.select ? 0 : h);
            const View V{&T, sc, &rings[w], &tsm[w], T.len1, T.len2, T.nb, T.nsteps, T.nwin, T.nstrips, h, lane, hd.max2[h], sc.vwords};
            AlignerOut &O = prm.outs[T.out[h]];
            BpList L{O.n_bp, bp_s[w]};
            __syncwarp();
            for (int i = lane; i < L.n * 4; i += 32) L.bp[i >> 2][i & 3] = O.bp[i >> 2][i & 3];
            __syncwarp();
            int alt = -1;
            const long long pc = clock64();
            for (int which = 0; which < 2; ++which) {
                int idx = 0, nl = 0, nr = 0;
                if (which == 0) {
                    nl = left_blocks(V, L.bp[0][0], L.bp[0][1], nullptr, 0, lane);
                    nr = right_blocks(V, L.bp[0][2], L.bp[0][3], nullptr, lane);
                } else {
                    for (idx = L.n - 1; idx >= 1; --idx) {
                        nl = left_blocks(V, L.bp[idx][0], L.bp[idx][1], nullptr, 0, lane);
                        nr = right_blocks(V, L.bp[idx][2], L.bp[idx][3], nullptr, lane);
                        if (nl + nr > 0) break;
                    }
                    if (idx < 1) { nl = nr = 0; idx = -1; }
                    alt = idx;
                }
                int off = 0;
                if (nl + nr > 0) {
                    if (lane == 0) off = atomicAdd(prm.pool_used, nl + nr);
                    off = __shfl_sync(0xFFFFFFFFu, off, 0);
                    if (off + nl + nr <= prm.pool_cap) {
                        left_blocks(V, L.bp[idx][0], L.bp[idx][1], prm.pool + off, nl, lane);
                        right_blocks(V, L.bp[idx][2], L.bp[idx][3], prm.pool + off + nl, lane);
                    } else if (lane == 0) O.status = -4;          // pool overflow: host retries with a larger pool
                }
                if (lane == 0) {
                    O.n_blk[2 * which] = nl; O.n_blk[2 * which + 1] = nr;
                    O.blk_off[2 * which] = off; O.blk_off[2 * which + 1] = off + nl;
                }
            }
            if (lane == 0) { O.alt_index = alt; PROF_ADD(3, clock64() - pc); }
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
    static const bool prof = getenv("AGE_P2_PROF") != nullptr;
    if (prof) { unsigned long long z[8] = {}; cudaMemcpyToSymbol(g_prof, z, sizeof(z)); }
    cudaMemsetAsync(p.next, 0, 8 * sizeof(int32_t), st);
    cudaMemsetAsync(p.queue_n, 0, sizeof(int32_t), st);
    const int nb = (p.n_pairs + 3) / 4;
    age_select_kernel<<<std::min(nb, sms * 8), SEL_WARPS * 32, 0, st>>>(p);
    age_presweep_kernel<<<sms * PRE_OCC, PRE_WARPS * 32, 0, st>>>(p, 0);      // windows along the split row / column (maxima trace)
    cudaMemsetAsync(p.queue_n, 0, sizeof(int32_t), st);
    age_enum_kernel<<<std::min(nb, sms * P2_OCC), P2_WARPS * 32, 0, st>>>(p);
    age_presweep_kernel<<<sms * PRE_OCC, PRE_WARPS * 32, 0, st>>>(p, 1);      // windows along the traceback paths (scores trace)
    age_blocks_kernel<<<std::min(nb, sms * P2_OCC), P2_WARPS * 32, 0, st>>>(p);
    if (prof) {
        unsigned long long z[8];
        int qn = 0;
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(z, g_prof, sizeof(z));
        cudaMemcpy(&qn, p.queue_n, sizeof(int), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[p2 prof] pre-swept windows %d, slowest pair index %llu\n", qn, z[0] & 0xFFFFFF); z[0] = (z[0] >> 24) << 10;
                fprintf(stderr, "[p2 prof] pairs %d | Mcycles: slowest pair %.1f ranges %.1f enumerate %.1f blocks %.1f | on-demand re-sweeps fwd %.1f rev %.1f | slowest ranges %.2f (hot windows %llu, len2 %llu) enumerate %.2f\n",
                p.n_pairs, z[0] / 1e6, z[1] / 1e6, z[2] / 1e6, z[3] / 1e6, z[4] / 1e6, z[5] / 1e6, (double)((z[6] >> 32) << 10) / 1e6, (z[6] >> 16) & 0xFFFF, z[6] & 0xFFFF, z[7] / 1e6);
    }
}

// ------------------------------------------------------------------------------------------------
// int16x2 SIMD issue-rate microbenchmark (denominator of bench.py's int_roofline)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
age_int16_peak_kernel(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[8];
    const uint32_t b = seed * 0x00010001u + threadIdx.x, c = (seed ^ 0x5555u) | 0x00010001u;
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = (threadIdx.x + k * 977u) * 0x00030001u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {                       // 8 independent chains, 6 SIMD / logic instructions each
            uint32_t x = __vmaxs2(a[k], b);
            x = __viaddmax_s16x2_relu(x, c, b);
            x = (x ^ c) & 0xFFFEFFFEu;
            x = __vadd2(x, b);
            x = __vimax3_s16x2(x, a[k], c);
            a[k] = x & 0x7FFF7FFFu;
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r ^= a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

} // namespace age

extern "C" int age_int16_peak(int device, double *giga_lane_instr_per_s)
{
    using namespace age;
    if (!giga_lane_instr_per_s || cudaSetDevice(device) != cudaSuccess) return -2;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    uint32_t *d = nullptr;
    if (cudaMalloc(&d, (size_t)blocks * threads * 4) != cudaSuccess) return -2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        age_int16_peak_kernel<<<blocks, threads>>>(d, iters, 3u + rep);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return -2; }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *giga_lane_instr_per_s = (double)blocks * threads * iters * 8 * 6 / (best * 1e-3) / 1e9;
    return 0;
}
