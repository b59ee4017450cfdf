// age_device.cuh -- device-side task descriptors shared by the kernels and the host engine.
//
// Vocabulary (follows the reference, src/AsmvarAGE/src/AGEaligner.cpp):
//   aligner    one single-mode AGEaligner run (INDEL / TDUP / INVL / INVR): a forward fill, a reverse
//              fill, their running maxima, breakpoint enumeration and the traceback.
//   pair       two aligners with identical (len1, len2, scoring) that share one warp: aligner A lives in the
//              low int16 of every packed word, aligner B in the high int16 (s16x2 SIMD).
//   sweep      one direction (forward K1+K3 or reverse K2+K4) of a pair.
//   strip      256 consecutive contig rows = 32 lanes x 8 rows held in registers.
//   block      W = 4 consecutive reference-window columns: what one lane advances per wavefront step.
//   step       lane p works on block t-p in step t (anti-diagonal wavefront over lanes).
//   window     32 consecutive steps of one strip: checkpoint interval, tile of the split-maximum table and the
//              unit of trace recomputation.
//
// Score representation inside the sweeps: every value is 2*score + tag in a signed 16-bit lane.
//   tag = 1 when the cell's own trace is HORIZONTAL/VERTICAL (AGEaligner.cpp:1006-1013: the gap
//   penalty for *leaving* a cell depends on that cell's trace).  max(2d, 2h+1, 2v+1, 0) then yields
//   the reference's tie order (gap beats diagonal on ties, negative clamps to 0 and clears the trace)
//   in one VIADDMNMX.S16x2.RELU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace age {

constexpr int RPT    = 8;          // contig rows per lane
constexpr int STRIP  = 32 * RPT;   // rows per strip
constexpr int W      = 4;          // window columns per lane per step
constexpr int CKS    = 32;         // steps per window
constexpr int CKW    = 38;         // checkpoint words per lane: H,G,M of 8 rows + 3 x 4 outputs + 2 carried
constexpr int MAX_BP = 100;        // AGEaligner.hh:94

// Storage formats of the D plane (PairTask::dfmt).  RAW: 8 packed 16-bit pairs per (lane, block column) = 2 uint4.
// OFF8: the forward maxima of consecutive rows differ by at most max(match, mismatch, 0) per row (every DP step adds at
// most that, and prefix maxima inherit the bound), so the 8 entries of a (lane, block column) are stored as the first one
// plus 8-bit offsets: per (lane, step) one uint4 with the 4 column bases and one uint4 of offset bytes per column,
// 5 uint4 instead of 8.  Used when max(match, mismatch, 0) <= 15 (offsets <= 2*15*7 = 210).
constexpr int DFMT_RAW = 0, DFMT_OFF8 = 1;
__host__ __device__ inline int d_planes(int dfmt) { return dfmt == DFMT_OFF8 ? 5 : 2 * W; }     // uint4 per (lane, step)

// trace codes, AGEaligner.hh:41-44
enum { T_STOP = 0, T_DIAG = 1, T_HORI = 2, T_VERT = 3 };

// mode flags, AGEaligner.hh:72-76
enum { F_INDEL = 0x01, F_TDUP = 0x02, F_INV = 0x04, F_INVR = 0x08, F_INVL = 0x10 };

struct PairTask {
    int32_t len1, len2, nstrips, P;      // P = nstrips * STRIP (padded rows)
    int32_t nb, nsteps, nwin;            // blocks = ceil(len1 / W), steps = nb + 31, windows = ceil(nsteps / CKS)
    int32_t flag[2];                     // single-mode flag of aligner A / B (0 = no aligner in that half)
    int32_t out[2];                      // index into the AlignerOut array (-1 = none)
    int32_t select;                      // 1: only the better-scoring half needs breakpoints (ties -> half 0):
                                         //    -both partners (age_align.cpp:273) or INVL/INVR of one -inv job (AGEaligner.cpp:151)
    uint32_t c0x, kabs, flip, gborder;   // packed per half; see the sweep
    uint32_t m2, mm2;                    // 2*match, 2*mismatch packed per half (special-column path)
    // inputs.  xblk[b] describes the 4 columns of logical block b (1..nb) in processing order:
    //   .x / .y = PRMT selectors of columns 0,1 / 2,3 (16 bit each); .z = per column byte: bit0 special, bits1-3 / 4-6 codes
    uint4 *xblk_f, *xblk_r;              // [nb + 2]
    uint2 *ytab;                         // [P] {Ta, Tb}: bytes = 2*S(x in ACGT, y_row) of aligner A / B
    uint8_t *ycode;                      // [P] low nibble = code of A's row, high nibble = B's
    // raw inputs of the two aligners (device copies of age_job::seq1 / seq2; null = no aligner in that half): the pack
    // kernel derives xblk / ytab / ycode from them (Scorer.hh:24-39, Sequence.hh:138-150,190-198)
    const char *raw1[2], *raw2[2];
    int16_t match[2], mismatch[2];
    // strip boundary rows, planes H, G, M: [(nstrips-1)][3][nb + 1] uint4 (4 columns of a block each)
    uint4 *bnd_f, *bnd_r;
    // Diagonal-major planes, [strip][step][...][lane]: what the 32 lanes touch in one step is contiguous, for the
    // forward and the reverse sweep alike (the anti-diagonal of one wavefront is the anti-diagonal of the other).
    uint32_t *D;                         // forward maxima shifted by one row and one column:
                                         //   RAW  [s][t][w][h][lane][4]: k = 4h+j -> 2*Fmax[256s + 8l + k][W(t-l-1) + w]
                                         //   = what the reverse cell (row+1, col+1) adds (K6, :545-550)
                                         //   OFF8 [s][t][5][lane][4]: plane 0 = entry k = 0 of the 4 columns, plane 1+w = offsets of
                                         //   column w, word j = bytes {A_2j, A_2j+1, B_2j, B_2j+1} (A / B = the halves of the pair)
    uint32_t *Dlast;                     // [P + 1]  2*Fmax[r][len1]
    uint32_t *Dtail;                     // [W*nb + 1] 2*Fmax[P][c] (only read when len2 == P)
    uint32_t *Rfirst;                    // [P + 2]  2*Rmax[r][1]
    uint32_t *tilemax;                   // [nstrips][nwin][32] max over (lane's 8 rows x one window) of Fmax + Rmax
    uint32_t *ckpt_f, *ckpt_r;           // [nstrips][nwin][CKW][32] lane state at the start of every window
    // pass-2 trace planes of the aligner(s) of this pair that get breakpoints (one set when `select`, else one per half)
    char *p2set; uint64_t p2stride;
    int32_t k7;                          // a TDUP / inversion aligner is present: sets also hold the reverse maxima plane
    int32_t dfmt;                        // DFMT_RAW / DFMT_OFF8
    int32_t wide;                        // 1: one aligner (half 0) in 32-bit words with the reference's u16 storage made explicit
                                         //    (Ar<true>, age_sweep.cuh): every scoring / length the 16-bit halves cannot hold
};

// Numbers of the printed alignment (bpoint 0) that the host would otherwise have to rebuild the gapped strings for:
// what AliFragment::countAligned counts per fragment, and the three "identity at / outside / inside breakpoints" runs of
// printAlignment (AGEaligner.cpp:308-369) on either sequence, in positions of the sequences as passed (age_counts).
enum { CALL_ALIGNED = 0, CALL_IDENTIC = 2, CALL_GAP = 4, CALL_AT1 = 6, CALL_AT2 = 9, CALL_OUT1 = 12, CALL_OUT2 = 13, CALL_IN1 = 14, CALL_IN2 = 15, CALL_N = 16 };

struct AlignerOut {
    int32_t status, score, n_bp, alt_index;   // n_bp = -1: score only (the half lost its selection)
    int32_t n_blk[4];                    // left, right, alt_left, alt_right
    int32_t blk_off[4];                  // offsets into the block pool (in blocks)
    int32_t call[CALL_N];
    int32_t bp[MAX_BP][4];
};

struct Block { int32_t start1, start2, length; };

// Trace planes of one aligner (pass 2), filled window by window by re-running the sweep from its checkpoints: the
// windows of the split row / column by the parallel pre-sweep kernel, everything else on demand.  A window of a plane is
// a PAGE taken from the batch's trace pool when it is first swept (a pair looks at a few dozen of its hundreds or
// thousands of windows): `ptab` maps (plane, window) to the page.  Pages, in 4 KB granules of the pool:
//   code planes FS, FM, RS (0 .. 2)   2 granules: [step in window][lane] uint2, 16 bits per block column, 2 bits per row
//   RM (3)                            3 granules: the code page + [step][lane] hot words (bit 8w+k = Fmax + Rmax == max)
//   reverse maxima Mp (4, K7 only)    32 granules: [step][column][half of rows][lane] uint4, as the raw D layout but unshifted
struct P2Scratch {
    int32_t *ptab;                       // [5][nwt] first granule of the page
    char *pool;                          // the batch's trace pool
    uint32_t pool_granules;              // its size; granules 0 .. 31 are the dump every sweep writes to once the pool is full
    uint32_t *pool_next;                 // granules handed out so far (beyond the dump); > pool_granules - 32: overflow, the host runs again
    uint32_t *valid;                     // [9][vwords] window bits: the four planes, Mp, queued for the pre-sweeps (FM, RM, FS, RS)
    int32_t *rowrange;                   // [2][len2 + 2]
    int32_t *hotlist;                    // [1 + nwt] number and indices (strip * nwin + window) of the reverse windows at the maximum
    int32_t vwords, nwt;
};
constexpr int P2_GRANULE = 4096, P2_DUMP_GRANULES = 32;
__host__ __device__ inline int p2_page_granules(int plane) { return plane < 3 ? 2 : plane == 3 ? 3 : 32; }

struct PairHdr { int32_t max2[2]; int32_t need[2]; };   // split maxima (x2) and which halves get breakpoints

struct Pass2Params {
    const PairTask *pairs;
    int32_t n_pairs;
    AlignerOut *outs;
    PairHdr *hdr;
    Block *pool; int32_t pool_cap; int32_t *pool_used;
    uint2 *queue; int32_t queue_cap; int32_t *queue_n;   // pre-sweep work items {pair, half | plane << 1 | strip << 4 | window << 16}
    int32_t *next;                       // work counters: select, pre-sweep 1, enumerate, blocks, pre-sweep 2
    char *tpool; uint32_t tpool_granules; uint32_t *tpool_next;   // trace pool of the batch (P2Scratch)
};

// layout of the per-aligner part of pass 2 (shared by host and device): page table, window bits, row ranges, hot list
__host__ __device__ inline size_t p2_up(size_t v) { return (v + 255) / 256 * 256; }
__host__ __device__ inline size_t p2_vwords(size_t nwt) { return (nwt + 31) / 32 + 1; }
__host__ __device__ inline size_t p2_set_bytes(size_t nwt, size_t len2)
{
    return p2_up(5 * nwt * 4) + p2_up(9 * p2_vwords(nwt) * 4) + p2_up(2 * (len2 + 2) * 4) + p2_up((nwt + 1) * 4);
}
// granules all pages of one aligner would take (the bound the pool never has to exceed)
__host__ __device__ inline size_t p2_full_granules(size_t nwt, int k7) { return nwt * (size_t)(2 + 2 + 2 + 3 + (k7 ? 32 : 0)); }

// One sub-batch on one device: pairs [0, n_wide) run in the 32-bit kernels, the rest in the packed 16-bit ones.
// counters: 32 ints, zeroed by launch_sweeps (0 / 16: pair counters of the packed / wide sweep kernel, 1: blocks used in
// the pool, 3: granules taken from the trace pool, the rest: work counters of the pass-2 kernels).
struct BatchLaunch {
    const PairTask *pairs; int32_t n_pairs, n_wide;
    int32_t nw, nw_wide, cs, cs_wide;                   // warps per CTA and CTAs per cluster that share a pair in the sweep (sweep_plan)
    AlignerOut *outs; Block *pool; int32_t pool_cap;
    int32_t *counters;
    PairHdr *hdr; uint2 *queue; int32_t queue_cap;
    char *tpool; uint32_t tpool_granules;               // trace pool (P2Scratch); its counter is counters[3]
};

// launchers (age_kernels.cu)
void launch_pack(const PairTask *d_pairs, int n_pairs, cudaStream_t st);
void sweep_plan(int n_pairs, int max_strips, int *nw, int *cs);
void launch_sweeps(const BatchLaunch &b, cudaStream_t st);
void launch_pass2(const BatchLaunch &b, cudaStream_t st);
int  launches_per_batch(const BatchLaunch &b);          // kernels the two calls above start

} // namespace age
