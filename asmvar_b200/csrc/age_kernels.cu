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
#include "age_pass2.cuh"
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdlib>
#include <cstdio>
#include <algorithm>
#include <vector>

namespace age {

constexpr int SWEEP_WARPS = 4;
#ifndef SWEEP_OCC
#define SWEEP_OCC 4
#endif

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
                        if (T.wide) {       // one aligner, 16-bit table entries {A, C | G, T}, sign-extended to 32 bits by the PRMT
                            if (a < 4) sel[w] = (uint32_t)(2 * a) | ((uint32_t)(2 * a + 1) << 4) | ((uint32_t)((2 * a + 1) | 8) << 8) | ((uint32_t)((2 * a + 1) | 8) << 12);
                            else { sel[w] = 0; info |= (1u | ((uint32_t)a << 1) | (5u << 4)) << (8 * w); }
                        }
                        else if (a < 4 && e < 4) sel[w] = (uint32_t)a | ((uint32_t)(a | 8) << 4) | ((uint32_t)(4 + e) << 8) | ((uint32_t)((4 + e) | 8) << 12);
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
                if (T.wide) {
                    if (k == 0)
                        for (int x = 0; x < 4; ++x) tab[x >> 1] |= ((uint32_t)(2 * (x == yc ? T.match[0] : T.mismatch[0])) & 0xFFFFu) << (16 * (x & 1));
                    continue;
                }
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

// The hot kernel: persistent, one warp per pair; the work counter hands out the pairs longest first.  The forward sweep of
// all strips streams D, the reverse sweep folds it back in; both leave a checkpoint per window and no trace.
// (Tried in round 2 and dropped, DESIGN.md section 4: pass 2 of a pair by the warp that swept it -- 98.6 ms per batch instead
// of 57.4, and even compiled in but not executed it costs the sweep 13 %; one CTA per four pairs with the batch cut into
// chunks on several streams so that pass 2 overlaps the next chunk's sweep -- 56-60 ms instead of 55.3.)
template <bool WIDE>
__global__ void __launch_bounds__(SWEEP_WARPS * 32, SWEEP_OCC)
age_sweep_kernel(const PairTask *__restrict__ pairs, int n_pairs, int32_t *counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SweepSmem *sm = reinterpret_cast<SweepSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    dstage_init(sm[w].dstage, lane);
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= n_pairs) break;
        const PairTask &T = pairs[item];
        sweep_pair_fast<+1, WIDE>(T, sm[w].rings, sm[w].dstage, lane);
        __threadfence_block();
        __syncwarp();
        sweep_pair_fast<-1, WIDE>(T, sm[w].rings, sm[w].dstage, lane);
        cp_async_wait_all();
        __syncwarp();
    }
}

// Long alignments / small batches: several warps share one pair and run its strips as a pipeline (warp g of the group owns
// strips g, g + G, ...; strip si trails strip si-1 by 32+ steps through the boundary rows in global memory).  The group is
// the NW warps of a CTA -- or, when even 16 warps per pair leave SMs without work (a few dozen long pairs), the warps of a
// thread-block cluster of 2 or 4 CTAs on neighbouring SMs: the strips' progress counters and the work item then live in the
// shared memory of the cluster's first CTA and the others reach them through distributed shared memory.
constexpr int MAX_STRIPS_LONG = 128;

template <bool WIDE, bool CLUSTER>
__global__ void __launch_bounds__(512, 1)
age_sweep_long_kernel(const PairTask *__restrict__ pairs, int n_pairs, int32_t *counter)
{
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SweepSmem *sm = reinterpret_cast<SweepSmem *>(smem_raw);
    __shared__ int prog_own[MAX_STRIPS_LONG];
    __shared__ int item_own;
    // one CTA per pair: plain shared memory and __syncthreads (a cluster of one through the cluster API measured 12 % slower);
    // a cluster per pair: counters and work item of the first CTA through distributed shared memory, cluster barriers
    int rank = 0, cs = 1;
    int *prog = prog_own, *item_p = &item_own;
    if (CLUSTER) {
        cg::cluster_group cluster = cg::this_cluster();
        rank = (int)cluster.block_rank(); cs = (int)cluster.num_blocks();
        prog = cluster.map_shared_rank(prog_own, 0);
        item_p = cluster.map_shared_rank(&item_own, 0);
    }
    auto sync_all = [&]() { if (CLUSTER) cg::this_cluster().sync(); else __syncthreads(); };
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int g = rank * nw + w, G = nw * cs;                // this warp in the group that shares the pair
    dstage_init(sm[w].dstage, lane);
    for (;;) {
        sync_all();
        if (rank == 0) {
            if (threadIdx.x == 0) item_own = atomicAdd(counter, 1);
            for (int i = threadIdx.x; i < MAX_STRIPS_LONG; i += blockDim.x) prog_own[i] = 0;
        }
        sync_all();
        const int item = *reinterpret_cast<volatile int *>(item_p);
        if (item >= n_pairs) break;
        const PairTask &T = pairs[item];
        sweep_pair_fast<+1, WIDE>(T, sm[w].rings, sm[w].dstage, lane, g, G, prog);
        __threadfence();
        sync_all();
        if (rank == 0) for (int i = threadIdx.x; i < MAX_STRIPS_LONG; i += blockDim.x) prog_own[i] = 0;
        sync_all();
        sweep_pair_fast<-1, WIDE>(T, sm[w].rings, sm[w].dstage, lane, g, G, prog);
        cp_async_wait_all();
    }
    if (CLUSTER) cg::this_cluster().sync();                  // nobody leaves while a neighbour may still read its shared memory
}

// Warps that share one pair in the sweep of a sub-batch.  From sms * 16 pairs on, one warp per pair fills the GPU (the
// persistent kernel).  Below that each pair is shared by nw = 2 .. 16 warps of one CTA, times cs = 1, 2 or 4 CTAs of a
// cluster (age_sweep_long_kernel): the combination is taken that keeps most of the SMs' warp slots busy -- pairs in flight
// x warps over the slots of the waves they need, times the share of a pair's strip rounds that all its warps work in; ties
// go to the smaller group, whose pipeline has fewer fill and drain steps.
void sweep_plan(int n_pairs, int max_strips, int *nw_out, int *cs_out)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    *nw_out = 1; *cs_out = 1;
    const int force_nw = getenv("AGE_LONG_NW") ? atoi(getenv("AGE_LONG_NW")) : 0;   // test knobs: 1 = one warp per pair, 2..16 = pipelined,
    const int force_cs = getenv("AGE_LONG_CS") ? atoi(getenv("AGE_LONG_CS")) : 0;   // clusters of 2 / 4 CTAs
    if (max_strips > MAX_STRIPS_LONG) return;
    if (force_nw >= 1 && force_nw <= 16) {
        *nw_out = force_nw;
        if (force_nw > 1 && force_cs >= 2 && force_cs <= 4) *cs_out = force_cs;       // (AGE_LONG_CS=1 or unset: one CTA per pair)
        return;
    }
    if (n_pairs <= 0 || n_pairs >= sms * 16) return;
    double best_score = 0;
    for (int cs = 1; cs <= 4; cs *= 2)
        for (int nw = 1; nw <= 16; nw *= 2) {
            const int G = nw * cs;
            if ((G > 1 && G / 2 >= std::max(1, max_strips)) || (cs > 1 && nw < 8)) continue;     // clusters only of large CTAs
            const int cap = std::max(1, sms * (16 / nw) / cs);                                     // pairs in flight
            const int waves = (n_pairs + cap - 1) / cap, rounds = (max_strips + G - 1) / G;
            const double score = (double)n_pairs / ((double)waves * cap) * ((double)max_strips / ((double)rounds * G));
            if (score > best_score * 1.0001) { best_score = score; *nw_out = nw; *cs_out = cs; }
        }
}

template <bool WIDE>
static void launch_sweeps_t(const PairTask *d_pairs, int n_pairs, int nw, int cs, int32_t *d_counter, cudaStream_t st)
{
    if (n_pairs <= 0) return;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (nw > 1 && cs == 1) {
        const int smem = nw * (int)sizeof(SweepSmem);
        cudaFuncSetAttribute(age_sweep_long_kernel<WIDE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        age_sweep_long_kernel<WIDE, false><<<std::min(n_pairs, sms * (16 / nw)), nw * 32, smem, st>>>(d_pairs, n_pairs, d_counter);
        return;
    }
    if (nw > 1) {
        const int smem = nw * (int)sizeof(SweepSmem);
        cudaFuncSetAttribute(age_sweep_long_kernel<WIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = (unsigned)cs; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.blockDim = dim3((unsigned)nw * 32); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        int groups = sms * (16 / nw) / cs;                                 // resident groups: persistent, the counter feeds them
        cfg.gridDim = dim3((unsigned)(sms * (16 / nw)));
        int fit = 0;
        if (cudaOccupancyMaxActiveClusters(&fit, age_sweep_long_kernel<WIDE, true>, &cfg) == cudaSuccess && fit > 0) groups = std::min(groups, fit);
        else cudaGetLastError();
        cfg.gridDim = dim3((unsigned)(std::min(n_pairs, std::max(1, groups)) * cs));
        cudaLaunchKernelEx(&cfg, age_sweep_long_kernel<WIDE, true>, d_pairs, n_pairs, d_counter);
        return;
    }
    const int blocks = std::min((n_pairs + SWEEP_WARPS - 1) / SWEEP_WARPS, sms * SWEEP_OCC);   // persistent: the work counter feeds the warps
    const int smem = SWEEP_WARPS * (int)sizeof(SweepSmem);
    cudaFuncSetAttribute(age_sweep_kernel<WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    age_sweep_kernel<WIDE><<<blocks, SWEEP_WARPS * 32, smem, st>>>(d_pairs, n_pairs, d_counter);
}

template <bool WIDE> static void launch_pass2_t(const Pass2Params &p, cudaStream_t st);

// The sweep kernels of a sub-batch (wide pairs first) on `st`
void launch_sweeps(const BatchLaunch &b, cudaStream_t st)
{
    cudaMemsetAsync(b.counters, 0, 32 * sizeof(int32_t), st);
    if (b.n_wide > 0) launch_sweeps_t<true>(b.pairs, b.n_wide, b.nw_wide, b.cs_wide, b.counters + 16, st);
    if (b.n_pairs > b.n_wide) launch_sweeps_t<false>(b.pairs + b.n_wide, b.n_pairs - b.n_wide, b.nw, b.cs, b.counters + 0, st);
}

// Pass 2 of a sub-batch on `st`: select -> pre-sweep -> enumerate -> pre-sweep -> blocks, per kind of pair
void launch_pass2(const BatchLaunch &b, cudaStream_t st)
{
    for (int wide = 1; wide >= 0; --wide) {
        const int first = wide ? 0 : b.n_wide, n = wide ? b.n_wide : b.n_pairs - b.n_wide;
        if (n <= 0) continue;
        int32_t *cnt = b.counters + (wide ? 16 : 0);
        Pass2Params p{};
        p.pairs = b.pairs + first; p.n_pairs = n; p.outs = b.outs;
        p.hdr = b.hdr + first;
        const int wq = b.n_wide ? (b.n_pairs > b.n_wide ? b.queue_cap / 4 : b.queue_cap) : 0;      // share of the work list for the wide pairs
        p.queue = b.queue + (wide ? 0 : wq); p.queue_cap = std::max(1, wide ? wq : b.queue_cap - wq); p.queue_n = cnt + 2;
        p.pool = b.pool; p.pool_cap = b.pool_cap; p.pool_used = b.counters + 1; p.next = cnt + 4;
        p.tpool = b.tpool; p.tpool_granules = b.tpool_granules; p.tpool_next = reinterpret_cast<uint32_t *>(b.counters + 3);
        if (wide) launch_pass2_t<true>(p, st); else launch_pass2_t<false>(p, st);
    }
}

int launches_per_batch(const BatchLaunch &b)
{
    const int kinds = (b.n_wide > 0) + (b.n_pairs > b.n_wide);
    return kinds * 6;                                       // sweep, select, 2 pre-sweeps, enumerate, blocks
}

// ------------------------------------------------------------------------------------------------
// pass 2: select -> pre-sweep -> enumerate -> pre-sweep -> blocks
// ------------------------------------------------------------------------------------------------
constexpr int P2_WARPS = 4;
#ifndef P2_OCC
#define P2_OCC 4
#endif

// ---- select: split maxima, -both / -inv winner, work list of the windows along the split row / column -------------
constexpr int SEL_WARPS = 4;

template <bool WIDE>
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
        int max2[2]; bool need[2];
        pair_select<WIDE>(T, prm.outs, lane, max2, need);
        if (lane == 0) prm.hdr[item] = PairHdr{{max2[0], max2[1]}, {need[0] ? 1 : 0, need[1] ? 1 : 0}};
        for (int h = 0; h < 2; ++h) {
            if (!need[h]) continue;
            const P2Scratch sc = carve_set(T, T.select ? 0 : h, prm);
            for (int i = lane; i < 9 * sc.vwords; i += 32) sc.valid[i] = 0;
            __syncwarp();
            if (T.flag[h] != F_INDEL) {
                // K7: the rows that tie at the maximum are scanned along their whole length (:485-534): queue every
                // window of their strips -- the forward maxima trace of row r, the reverse one (with the maxima
                // themselves) of row r+1
                uint32_t *qf = sc.valid + 5 * sc.vwords, *qr = sc.valid + 6 * sc.vwords;
                auto pick = [&](uint32_t v) { return WIDE ? (int)v : (int)((v >> (16 * h)) & 0xFFFFu); };
                for (int rb = 0; rb <= len2; rb += 32) {
                    const int r = rb + lane;
                    bool tie = false;
                    if (r <= len2) {
                        const int a = r >= 1 ? pick(T.Dlast[r]) : 0, b = r + 1 <= len2 ? pick(T.Rfirst[r + 1]) : 0;
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
            const int nhot = build_hotlist<WIDE>(T, sc, h, max2[h], lane);
            for (int hi = 0; hi < nhot; ++hi) {
                if (lane >= 3) continue;
                const int wi = sc.hotlist[1 + hi];
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
        }
    }
}

// ---- pre-sweep: one warp per queued window, at full occupancy ------------------------------------------------------
constexpr int PRE_WARPS = 4;
#ifndef PRE_OCC
#define PRE_OCC 4
#endif

template <bool WIDE>
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
        const P2Scratch sc = carve_set(T, T.select ? 0 : h, prm);
        const bool with_mp = plane == 5;                     // reverse maxima trace + the maxima themselves (K7)
        const int pl = plane == 5 ? 3 : plane;
        // the window's pages come out of the trace pool (same rule as View::take_page)
        auto take = [&](int plane_) {
            uint32_t g = 0;
            if (lane == 0) {
                const uint32_t n = (uint32_t)p2_page_granules(plane_);
                g = atomicAdd(sc.pool_next, n) + P2_DUMP_GRANULES;
                if (g + n > sc.pool_granules) g = 0;
                sc.ptab[plane_ * sc.nwt + s * T.nwin + win] = (int32_t)g;
            }
            g = __shfl_sync(0xFFFFFFFFu, g, 0);
            return sc.pool + (size_t)g * P2_GRANULE;
        };
        char *pg = take(pl);
        TraceOut TO{reinterpret_cast<uint2 *>(pg), reinterpret_cast<uint32_t *>(pg + 2 * P2_GRANULE), nullptr, WIDE ? 0 : 16 * h, prm.hdr[q.x].max2[h]};
        if (with_mp) TO.Mp = reinterpret_cast<uint32_t *>(take(4));
        switch (pl) {
        case 0:  sweep_window_trace<+1, MODE_TRACE_S, WIDE>(T, rings[w], tsm[w], lane, s, win, TO); break;
        case 1:  sweep_window_trace<+1, MODE_TRACE_M, WIDE>(T, rings[w], tsm[w], lane, s, win, TO); break;
        case 2:  sweep_window_trace<-1, MODE_TRACE_S, WIDE>(T, rings[w], tsm[w], lane, s, win, TO); break;
        default: sweep_window_trace<-1, MODE_TRACE_M, WIDE>(T, rings[w], tsm[w], lane, s, win, TO); break;
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
template <bool WIDE>
__device__ void queue_path(const Pass2Params &prm, const View<WIDE> &V, int item, int plane, int r0, int c0, int dir, int steps)
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
template <bool WIDE>
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
        const PairHdr hd = prm.hdr[item];
        for (int h = 0; h < 2; ++h) {
            if (!hd.need[h]) continue;
            const P2Scratch sc = carve_set(T, T.select ? 0 : h, prm);
#ifdef AGE_P2_PROF
            __shared__ unsigned int od_s[P2_WARPS][8];
            if (lane < 8) od_s[w][lane] = 0;
            __syncwarp();
            const View<WIDE> V{&T, sc, &rings[w], &tsm[w], T.len1, T.len2, T.nb, T.nsteps, T.nwin, T.nstrips, h, lane, hd.max2[h], sc.vwords, od_s[w]};
#else
            const View<WIDE> V{&T, sc, &rings[w], &tsm[w], T.len1, T.len2, T.nb, T.nsteps, T.nwin, T.nstrips, h, lane, hd.max2[h], sc.vwords};
#endif
            AlignerOut &O = prm.outs[T.out[h]];
            BpList L{0, bp_s[w]};
            P2_PROF_T0(pc_pair);
            enumerate_aligner(V, T.flag[h], L, O, lane);
            P2_PROF_PAIR(0, item, pc_pair);
            P2_PROF_RESET(pc_pair);
            // the tracebacks of bpoint 0 and of the alternative printAlignment shows (the last one, normally)
            for (int which = 0; which < 2; ++which) {
                const int idx = which == 0 ? 0 : L.n - 1;
                if (which == 1 && idx < 1) break;
                const int lg1 = L.bp[idx][0], lg2 = L.bp[idx][1], rg1 = L.bp[idx][2], rg2 = L.bp[idx][3];
                const int fl = V.F2(lg2, lg1) >> 1, fr = max((hd.max2[h] >> 1) - fl, 0);     // scores of the two fragments
                queue_path(prm, V, item, 0, lg2, lg1, -1, fl + fl / 4 + 64);
                queue_path(prm, V, item, 2, rg2, rg1, +1, fr + fr / 4 + 64);
            }
            P2_PROF_PAIR(1, item, pc_pair);
#ifdef AGE_P2_PROF
            if (lane == 0 && item < PROF_PAIRS) { g_pair_cycles[3][item] += od_s[w][0]; g_pair_cycles[5][item] += od_s[w][1]; g_pair_cycles[6][item] += od_s[w][2]; g_pair_cycles[7][item] += od_s[w][3]; for (int q = 4; q < 8; ++q) g_pair_cycles[4 + q][item] += od_s[w][q]; }
#endif
            __syncwarp();
        }
    }
}

// ---- blocks: K10-K12 tracebacks of bpoint 0 and of the alternative shown by printAlignment (:287-306) ---------------------
// (the walks are executed by the whole warp: a trace window that is not resident is re-swept by all lanes)
template <bool WIDE>
__global__ void __launch_bounds__(P2_WARPS * 32, P2_OCC)
age_blocks_kernel(Pass2Params prm)
{
    __shared__ int bp_s[P2_WARPS][MAX_BP][4];
    __shared__ WarpRings rings[P2_WARPS];
    __shared__ TraceSmem tsm[P2_WARPS];
    __shared__ Block blk_s[P2_WARPS][2 * BLK_STAGE];      // blocks of the traceback at hand (emit_blocks)
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
            const P2Scratch sc = carve_set(T, T.select ? 0 : h, prm);
#ifdef AGE_P2_PROF
            __shared__ unsigned int od_s[P2_WARPS][8];
            if (lane < 8) od_s[w][lane] = 0;
            __syncwarp();
            const View<WIDE> V{&T, sc, &rings[w], &tsm[w], T.len1, T.len2, T.nb, T.nsteps, T.nwin, T.nstrips, h, lane, hd.max2[h], sc.vwords, od_s[w]};
#else
            const View<WIDE> V{&T, sc, &rings[w], &tsm[w], T.len1, T.len2, T.nb, T.nsteps, T.nwin, T.nstrips, h, lane, hd.max2[h], sc.vwords};
#endif
            AlignerOut &O = prm.outs[T.out[h]];
            BpList L{O.n_bp, bp_s[w]};
            __syncwarp();
            for (int i = lane; i < L.n * 4; i += 32) L.bp[i >> 2][i & 3] = O.bp[i >> 2][i & 3];
            __syncwarp();
            P2_PROF_T0(pc_pair);
            emit_blocks(V, T.flag[h], L, O, prm.pool, prm.pool_cap, prm.pool_used, lane, blk_s[w]);
            P2_PROF_PAIR(2, item, pc_pair);
#ifdef AGE_P2_PROF
            if (lane == 0 && item < PROF_PAIRS) g_pair_cycles[4][item] += od_s[w][0];
#endif
        }
    }
}

template <bool WIDE>
static void launch_pass2_t(const Pass2Params &p, cudaStream_t st)
{
    if (p.n_pairs <= 0) return;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nb = (p.n_pairs + 3) / 4;
    age_select_kernel<WIDE><<<std::min(nb, sms * 8), SEL_WARPS * 32, 0, st>>>(p);
    age_presweep_kernel<WIDE><<<sms * PRE_OCC, PRE_WARPS * 32, 0, st>>>(p, 0);      // windows along the split row / column (maxima trace)
    cudaMemsetAsync(p.queue_n, 0, sizeof(int32_t), st);
    age_enum_kernel<WIDE><<<std::min(nb, sms * P2_OCC), P2_WARPS * 32, 0, st>>>(p);
    age_presweep_kernel<WIDE><<<sms * PRE_OCC, PRE_WARPS * 32, 0, st>>>(p, 1);      // windows along the traceback paths (scores trace)
    age_blocks_kernel<WIDE><<<std::min(nb, sms * P2_OCC), P2_WARPS * 32, 0, st>>>(p);
#ifdef AGE_P2_PROF
    {   // cycle accounting of pass 2 (diagnostic builds only): summed over the warps of the launch
        unsigned long long z[16];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(z, g_prof, sizeof(z));
        fprintf(stderr, "[p2 prof] pairs %d | Mcycles summed over warps: ranges / K7 forward %.1f, enumerate / K7 reverse %.1f, blocks %.1f, "
                        "on-demand re-sweeps forward %.1f reverse %.1f (counts %llu / %llu)\n",
                p.n_pairs, z[1] / 1e6, z[2] / 1e6, z[3] / 1e6, z[4] / 1e6, z[5] / 1e6, z[6], z[7]);
        fprintf(stderr, "[p2 prof] on-demand windows per plane FS %llu FM %llu RS %llu RM %llu, with maxima RM %llu\n", z[8], z[9], z[10], z[11], z[15]);
        unsigned long long zero[16] = {};
        cudaMemcpyToSymbol(g_prof, zero, sizeof(zero));
        unsigned long long z2[8];
        cudaMemcpyFromSymbol(z2, g_prof2, sizeof(z2));
        if (z2[0]) fprintf(stderr, "[p2 prof] K7: %llu tie rows, %llu forward / %llu reverse start cells taking %.1f / %.1f Mcycles, ensure_rm %.1f Mcycles\n",
                           z2[0], z2[1], z2[2], z2[3] / 1e6, z2[4] / 1e6, z2[5] / 1e6);
        cudaMemcpyToSymbol(g_prof2, zero, sizeof(z2));
        // the slowest pairs of the launch (the enumerate / blocks kernels last as long as their slowest warp)
        const int n = std::min(p.n_pairs, (int)PROF_PAIRS);
        std::vector<unsigned int> cyc(12 * (size_t)PROF_PAIRS);
        cudaMemcpyFromSymbol(cyc.data(), g_pair_cycles, cyc.size() * sizeof(unsigned int));
        std::vector<PairTask> pt(n);
        cudaMemcpy(pt.data(), p.pairs, n * sizeof(PairTask), cudaMemcpyDeviceToHost);
        std::vector<int> order(n);
        for (int i = 0; i < n; ++i) order[i] = i;
        auto tot = [&](int i) { return (unsigned long long)cyc[i] + cyc[PROF_PAIRS + i]; };
        std::sort(order.begin(), order.end(), [&](int a, int b) { return tot(a) > tot(b); });
        unsigned long long sum = 0, od_e = 0, od_b = 0;
        for (int i = 0; i < n; ++i) { sum += tot(i); od_e += cyc[3 * PROF_PAIRS + i]; od_b += cyc[4 * PROF_PAIRS + i]; }
        fprintf(stderr, "[p2 prof] enumerate kernel, cycles per pair: mean %.2f M, median %.2f M, p90 %.2f M, p99 %.2f M; windows re-swept on demand: %llu enumerate, %llu blocks\n",
                sum / 1e6 / std::max(n, 1), n ? tot(order[n / 2]) / 1e6 : 0.0, n ? tot(order[n / 10]) / 1e6 : 0.0, n ? tot(order[n / 100]) / 1e6 : 0.0, od_e, od_b);
        for (int k = 0; k < std::min(n, 12); ++k) {
            const int i = order[k];
            AlignerOut ao[2] = {};
            for (int h = 0; h < 2; ++h) if (pt[i].out[h] >= 0) cudaMemcpy(&ao[h], p.outs + pt[i].out[h], 32, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[p2 prof]   pair %5d: len1 %d len2 %d, enumerate %.2f M + path queueing %.2f M, blocks %.2f M cycles; K6: %u hot windows, ranges (K7: start cells) %.2f M, forward scan (K7: ensure_rm) %.2f M, %u narrow passes (K7: tie rows), %u wide rows (K7: start cells), %u walks taking %.2f M; on demand %u + %u windows; n_bp %d / %d, score %d / %d\n",
                    i, pt[i].len1, pt[i].len2, cyc[i] / 1e6, cyc[PROF_PAIRS + i] / 1e6, cyc[2 * PROF_PAIRS + i] / 1e6, cyc[6 * PROF_PAIRS + i], cyc[5 * PROF_PAIRS + i] / 1e6, cyc[7 * PROF_PAIRS + i] / 1e6, cyc[8 * PROF_PAIRS + i], cyc[9 * PROF_PAIRS + i], cyc[10 * PROF_PAIRS + i], cyc[11 * PROF_PAIRS + i] / 1e6,
                    cyc[3 * PROF_PAIRS + i], cyc[4 * PROF_PAIRS + i],
                    ao[0].n_bp, ao[1].n_bp, ao[0].score, ao[1].score);
        }
        std::fill(cyc.begin(), cyc.end(), 0u);
        cudaMemcpyToSymbol(g_pair_cycles, cyc.data(), cyc.size() * sizeof(unsigned int));
    }
#endif
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
